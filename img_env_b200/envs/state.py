"""ImageState: the nine-field observation container of the reference (envs/state/state.py:4-28) with the same constructor
order, attribute names, `len()` and printable form.  The fields are torch CUDA tensors straight from the library (or numpy
arrays when the env is built with numpy_state=True); all of them have one row per robot."""


class ImageState:
    FIELDS = ("vector_states", "sensor_maps", "is_collisions", "is_arrives", "lasers", "ped_vector_states", "ped_maps",
              "step_ds", "ped_min_dists")
    __slots__ = FIELDS

    def __init__(self, *values, **named):
        given = dict(zip(self.FIELDS, values))
        overlap = set(given) & set(named)
        if overlap:
            raise TypeError("ImageState got multiple values for %s" % sorted(overlap))
        given.update(named)
        missing = [f for f in self.FIELDS if f not in given]
        if missing or len(given) != len(self.FIELDS):
            raise TypeError("ImageState takes exactly the fields %s (missing %s)" % (self.FIELDS, missing))
        rows = {len(v) for v in given.values()}
        assert len(rows) == 1, "every ImageState field needs one row per robot, got row counts %s" % sorted(rows)
        for f in self.FIELDS:
            setattr(self, f, given[f])

    def __len__(self):
        return len(self.vector_states)

    def __str__(self):
        return "Image State Info:\n" + "\n".join("        %s: %s" % (f, getattr(self, f)) for f in self.FIELDS)
