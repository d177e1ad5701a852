"""ImageState: the nine-field observation container of the reference (envs/state/state.py:4-28), unchanged
API; fields are torch CUDA tensors (or numpy arrays when the env is built with numpy_state=True)."""


class ImageState:
    FIELDS = ("vector_states", "sensor_maps", "is_collisions", "is_arrives", "lasers", "ped_vector_states", "ped_maps",
              "step_ds", "ped_min_dists")

    def __init__(self, vector_states, sensor_maps, is_collisions, is_arrives, lasers, ped_vector_states, ped_maps, step_ds,
                 ped_min_dists):
        assert len(vector_states) == len(sensor_maps) == len(is_collisions) == len(is_arrives) == len(lasers) \
            == len(ped_vector_states) == len(ped_maps) == len(step_ds) == len(ped_min_dists)
        self.vector_states = vector_states
        self.sensor_maps = sensor_maps
        self.is_collisions = is_collisions
        self.is_arrives = is_arrives
        self.lasers = lasers
        self.ped_vector_states = ped_vector_states
        self.ped_maps = ped_maps
        self.ped_min_dists = ped_min_dists
        self.step_ds = step_ds

    def __len__(self):
        return len(self.vector_states)

    def __str__(self):
        return "Image State Info:\n" + "\n".join("        %s: %s" % (f, getattr(self, f)) for f in self.FIELDS)
