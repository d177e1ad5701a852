// Observation stage: collision codes, egocentric view raster, laser rays, laser_map reconstruction, 400->48 cubic
// resize, pedestrian observation packing.
// Reference: ImgEnv::view_ped/view_robot (img_env.cpp:594-674), Agent::draw (agent.cpp:285-327),
// PedAgent::draw_leg (agent.cpp:737-774), Agent::view (agent.cpp:356-509), Agent::bresenhamLine
// (agent.cpp:511-624), ImgEnv::get_states (img_env.cpp:547-587) and the Python post-processing
// yaml_env.py:392-481.  See DESIGN.md for how each phase maps to the reference and why the
// results are identical.
#pragma once
#include "state.cuh"
#include "kin.cuh"
#include "foot.cuh"

#ifndef VIEW_THREADS
#define VIEW_THREADS 256
#endif
#ifndef VIEW_STATS
#define VIEW_STATS 0          // instrumented build (tools/view_stats.py): work counters per robot in Dev::counters[4..]
#endif
#ifndef VIEW_MIN_CTAS
#define VIEW_MIN_CTAS 4
#endif
//               // 64 registers per thread: the register file holds 32 warps per SM either way
#define FX_GUARD 8192u                 // |frac - 0.5| below 2^-19 cells -> exact fp64 fallback

struct ViewShared {
    ViewConst k;             // pose-dependent constants (foot.cuh), bulk-copied from d.vconst
    unsigned long long bar[2];      // mbarriers: [0] constants landed, [1] static tables landed
    int red[8];              // counters / per-warp partial sums
    int coll_key;
    int n_cnear, n_dirty, n_seglist;
    unsigned long long near_pack;   // number of near records << 32 | their words so far (one atomic hands out slot and word offset)
    int hmin[64];            // per block of rays: smallest hit step (Chebyshev distance of the hit cell), NOHIT >> 22 if none
    int stat[4];             // VIEW_STATS only: candidates, raster cells pushed, all-shadow outputs
    int hmax[64];            // per block of rays: largest hit step, NOHIT >> 22 (1023) when a ray of the block has no hit
};

// python float floor division (Objects/floatobject.c float_floor_div) used by yaml_env.py:414-415
__device__ __forceinline__ double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod) { if ((wx < 0) != (mod < 0)) { div -= 1.0; } }
    double floordiv;
    if (div) { floordiv = floor(div); if (div - floordiv > 0.5) floordiv += 1.0; }
    else floordiv = copysign(0.0, vx / wx);
    return floordiv;
}

// does ray k touch view cell (pr,pc)?  Closed form of the integer line walk in
// agent.cpp:517-622: with w=|x2-x1|, h=|y2-y1| the i-th visited cell is
//   w > h : (x1 + dx*i, y1 + dy*floor((2*h*i + w) / (2*w))),  i < w
//   else  : (x1 + dx*floor((2*w*i + h) / (2*h)), y1 + dy*i),  i < h
// Returns the step index i or -1.
__device__ __forceinline__ int ray_touch(int ox, int oy, int ex, int ey, int pr, int pc) {
    int w = ex - ox, h = ey - oy;
    int dx = w > 0 ? 1 : -1, dy = h > 0 ? 1 : -1;
    w = abs(w); h = abs(h);
    int a = pr - ox, b = pc - oy;
    if (w > h) {
        int i = abs(a);
        if (i >= w || a != dx * i) return -1;
        int m = abs(b);
        if (b != dy * m) return -1;
        int num = 2 * h * i + w;
        if (num < 2 * w * m || num >= 2 * w * (m + 1)) return -1;
        return i;
    } else {
        int i = abs(b);
        if (i >= h || b != dy * i) return -1;
        int m = abs(a);
        if (a != dx * m) return -1;
        int num = 2 * w * i + h;
        if (num < 2 * h * m || num >= 2 * h * (m + 1)) return -1;
        return i;
    }
}

// position of the n-th (0-based) set bit of m; m has more than n bits set.  (The generic __fns intrinsic costs ~55 instructions.)
__device__ __forceinline__ int nth_set_bit(unsigned m, int n) {
    int pos = 0, c;
    c = __popc(m & 0xFFFFu); if (n >= c) { n -= c; pos = 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (n >= c) { n -= c; pos += 8; m >>= 8; }
    c = __popc(m & 0xFu);    if (n >= c) { n -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u);    if (n >= c) { n -= c; pos += 2; m >>= 2; }
    return pos + (n >= (int)(m & 1u) ? 1 : 0);
}

// Rare continuation of the laser_map pixel rule (phase D): the pixel lies behind its top ray's hit and was not shadow
// written by it, so the rays below (kh-1 .. kl) decide, highest first.  Out of line to keep the kernel's hot loop small.
__device__ __noinline__ unsigned pixel_code_below(const unsigned* hitkey, const short* rend, int ox, int oy, int pr, int pc, int kh, int kl) {
    for (int kk = kh - 1; kk >= kl; kk--) {
        const int i = ray_touch(ox, oy, rend[2 * kk], rend[2 * kk + 1], pr, pc);
        if (i < 0) continue;
        const unsigned key2 = hitkey[kk];
        const int hp2 = (int)(key2 >> 22);
        if (i < hp2) return 3u;
        if (i == hp2) return 0u;
        if (pr != (int)((key2 >> 11) & 2047) && pc != (int)(key2 & 2047)) return 2u;
    }
    return 2u;
}
// Ray-hit candidates of one raster cell, resolved on the spot (only when the shared-memory lists are full).
__device__ __noinline__ void cell_rays_inline(unsigned* hitkey, const short* rend, int ox, int oy, int pr, int pc, int kl, int kh) {
    for (int k = kl; k <= kh; k++) {
        const int i = ray_touch(ox, oy, rend[2 * k], rend[2 * k + 1], pr, pc);
        if (i >= 0) atomicMin(&hitkey[k], ((unsigned)i << 22) | ((unsigned)pr << 11) | (unsigned)pc);
    }
}

// Shared-memory plan of k_view (bytes for the canonical 400x400 view / 1000 rays / 48x48 outputs; ~50 KB -> 4 CTAs per SM):
//   region A  occupancy raster 400*13*4 = 20.8 KB (x2 with lasers off: + the "known" plane)            phases B-D
//   region B  raster cells crossed by many rays BL2_CAP*4 = 2 KB (the others update their rays on the spot)  phases B-C
//             later: list of output pixels that need the full evaluation img*img*2 = 4.6 KB + hit prefix counts 2 KB   phases D-F
//   hitkey    range_total*4 (first hit per ray), ray end cells range_total*4, needed-line indices, FOV spans vh*8,
//             list of world blocks under the FOV that hold static candidates INV_MAX_BLOCKS*4,
//             footprint records near the robot: NP*2 (part ids) + (NP+1)*4 (running word counts) + colliding parts
//   (no 400x400 pixel buffer: the laser_map values are evaluated per output pixel of the resize)
#define BL2_CAP 512          // cells touched by many rays (close to the origin): listed, then processed warp-cooperatively
#ifndef BL_HEAVY
#define BL_HEAVY 12
#endif
#define NOHIT 0xFFFFFFFFu
#define CN_CAP 32            // footprint records that overlap the observer's own footprint box (collision candidates)
#define NEAR_CACHE 64        // near records whose header and word offset are kept in shared memory for phase B
#define CAND_CHUNK (2 * VIEW_THREADS)   // candidate words per pass of phase B (two per thread)
#define NEAR_ALL 0x8000u     // near-list flag: read the part's occupancy words (it may touch the FOV edge), not its candidates
#define ET_SHIFT 4           // edge tiles are 16x16 view pixels

#define INV_EPS 0.004f       // band around a cell edge inside which the inverse rasterisation runs the exact forward map
#define INV_MAX_BLOCKS 512   // 32x32-cell world blocks under the FOV; more -> forward (tile) rasterisation
inline ViewLayout view_layout(const Cfg& c) {
    ViewLayout L;
    size_t off = 0;
    L.sh = off; off += (sizeof(ViewShared) + 15) & ~(size_t)15;
    size_t occ = c.inverse_ok ? 0 : (size_t)c.vh * c.vwb * 4 * (c.use_laser ? 1 : 2);   // no raster at all in the world->view mode
    L.regA = off; off += (occ + 15) & ~(size_t)15;
    const size_t dl = ((size_t)c.img * c.img * 2 + 15) & ~(size_t)15;
    size_t bl = (size_t)BL2_CAP * 4, hb = dl + ((size_t)c.range_total + 1) * 2;
    if (!c.inverse_ok) bl = bl > (size_t)((c.vh + 31) / 32) * c.vwb * 2 ? bl : (size_t)((c.vh + 31) / 32) * c.vwb * 2;   // forward mode: tile list
    L.regB = off; L.hpre = off + dl; off += ((bl > hb ? bl : hb) + 15) & ~(size_t)15;
    L.hitkey = off; off += ((size_t)c.range_total * 4 + 15) & ~(size_t)15;
    L.rays = off; off += ((size_t)c.range_total * 4 + 15) & ~(size_t)15;
    L.need = off; off += ((size_t)c.ns * 2 + 15) & ~(size_t)15;
    L.spans = off; off += ((size_t)c.vh * 8 + 15) & ~(size_t)15;
    L.blocks = off; off += (size_t)INV_MAX_BLOCKS * 4;
    L.near = off; off += ((size_t)c.NP * 2 + 15) & ~(size_t)15;
    L.npre = off; off += ((size_t)(c.NP + 1) * 4 + 15) & ~(size_t)15;
    L.chdr = off; off += (size_t)CN_CAP * 16;
    L.coff = off; off += (size_t)CN_CAP * 4;
    L.nhdr = off; off += (size_t)NEAR_CACHE * 16;
    L.noff = off; off += (size_t)NEAR_CACHE * 4;
    L.cword = off; off += (size_t)CAND_CHUNK * 4;
    L.cmeta = off; off += (size_t)CAND_CHUNK * 4;
    L.cpre = off; off += (size_t)CAND_CHUNK * 2;
    L.cwsum = off; off += (size_t)(VIEW_THREADS / 32) * 4;
    L.seglist = off; off += (((size_t)c.img * c.img + 7) / 8 * 2 + 15) & ~(size_t)15;
    L.total = off + 16;
    return L;
}
inline size_t view_smem_bytes(const Cfg& c) { return view_layout(c).total; }

__device__ __forceinline__ unsigned hit_key(int i, int x, int y) { return ((unsigned)i << 22) | ((unsigned)x << 11) | (unsigned)y; }

// exact reference operation sequence for a view pixel's world cell (map2world, tf multiply, world2map): only inside the
// fixed-point form's guard band, hence out of line
__device__ __noinline__ void exact_cell(const Tf2& view_world, double res, int i, int j, int& cx, int& cy) {
    double wx, wy;
    tf_apply(view_world, i * res, j * res, wx, wy);
    cx = world2cell(wx, res); cy = world2cell(wy, res);
}
// Collision code and done flag of one robot, once its code of this step is known (img_env.cpp:547-587, yaml_env.py:316, 374-376);
// the rest of the state vector is written by k_view_consts.
__device__ __forceinline__ void view_publish_code(const Dev& d, int idx, int code, int is_reset) {
    const int arr = RBF(d, RB_ARR, idx) != 0.0;
    d.o_coll[idx] = (int8_t)code;
    RBF(d, RB_DONE, idx) = is_reset ? 0.0 : (double)min(1, min(code, 1) + arr);   // yaml_env.py:316, 374-376
}

// FWD = false: lasers on and a FOV of ordinary size -> world->view rasterisation, no raster in shared memory (the hot variant);
// FWD = true: lasers off (the "known" plane needs every FOV pixel) or a huge FOV -> forward rasterisation into a raster.
// VIEW_STATS builds: cycles between phase boundaries of a CTA (thread 0's clock), summed into Dev::counters[20 + phase]
#if VIEW_STATS
#define VIEW_MARK(ph) do { if (tid == 0) { const long long t_now = clock64(); atomicAdd(d.counters + 20 + (ph), (unsigned long long)(t_now - t_mark)); t_mark = t_now; } } while (0)
#else
#define VIEW_MARK(ph) do { } while (0)
#endif
// MINB = CTAs per SM the kernel is compiled for: 4 (64 registers) or 5 (48 registers, a few more spills, 25 % more warps in
// flight).  Measured: 5 wins on scenes with few parts (C1 +2 %, C5 +5 %), 4 on C4's 800 parts per scene (+1.2 %); imgenv.cu picks.
template <bool DEBUG_FULL, bool FWD, int MINB = VIEW_MIN_CTAS>
__global__ void __launch_bounds__(VIEW_THREADS, MINB) k_view(Dev d, const int* scene_ids, int is_reset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Cfg& c = d.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_mark = VIEW_STATS ? clock64() : 0; (void)t_mark;
    const int sl = blockIdx.x / c.R, r = blockIdx.x % c.R;
    if (d.n_dev && sl >= *d.n_dev) return;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int idx = s * c.R + r;
    const RobotType ty = d.types[d.type_of[r]];
    const ViewLayout& L = d.vl;          // (computed once on the host)
    const int vwb = c.vwb, vh = c.vh, vw = c.vw;

    ViewShared* sh = reinterpret_cast<ViewShared*>(smem_raw + L.sh);
    uint32_t* occ = reinterpret_cast<uint32_t*>(smem_raw + L.regA);
    uint32_t* known = occ + (size_t)vh * vwb;                               // only when !use_laser
    uint32_t* blist2 = reinterpret_cast<uint32_t*>(smem_raw + L.regB);
    unsigned short* dirty = reinterpret_cast<unsigned short*>(smem_raw + L.regB);   // after phase C
    unsigned* hbits = reinterpret_cast<unsigned*>(smem_raw + L.hpre);              // after phase C: bit k = ray k hit something
    short* spans = reinterpret_cast<short*>(smem_raw + L.spans);
    uint32_t* blocks = reinterpret_cast<uint32_t*>(smem_raw + L.blocks);
    unsigned* hitkey = reinterpret_cast<unsigned*>(smem_raw + L.hitkey);
    short* rend = reinterpret_cast<short*>(smem_raw + L.rays);
    short* need = reinterpret_cast<short*>(smem_raw + L.need);
    unsigned short* near = reinterpret_cast<unsigned short*>(smem_raw + L.near);
    unsigned* npre = reinterpret_cast<unsigned*>(smem_raw + L.npre);
    int4* chdr = reinterpret_cast<int4*>(smem_raw + L.chdr);
    int* coff = reinterpret_cast<int*>(smem_raw + L.coff);
    int4* nhdr = reinterpret_cast<int4*>(smem_raw + L.nhdr);
    int* noff = reinterpret_cast<int*>(smem_raw + L.noff);
    uint32_t* cword = reinterpret_cast<uint32_t*>(smem_raw + L.cword);
    uint32_t* cmeta = reinterpret_cast<uint32_t*>(smem_raw + L.cmeta);
    unsigned short* cpre = reinterpret_cast<unsigned short*>(smem_raw + L.cpre);
    int* cwsum = reinterpret_cast<int*>(smem_raw + L.cwsum);
    unsigned short* seglist = reinterpret_cast<unsigned short*>(smem_raw + L.seglist);   // phase D: segments of 8 outputs that need a closer look
    const int npx = c.img * c.img;
    // [npx] u32: lowest | highest top ray over the source pixels of an output | farthest source pixel, then [npx] u16: the
    // output's float16 value when none of those rays hits anything
    const uint32_t* okk = d.ostat + (size_t)ty.ostat_off;
    const uint16_t* oval = reinterpret_cast<const uint16_t*>(okk + npx);
    const short* ctap = d.cubic_tap;
    const short* ccoef = d.cubic_coef;
    const uint16_t* lut16 = d.f16_lut;

    const int4* fhdr = d.foot_hdr + (size_t)s * c.NP;
    const uint32_t* fwords = d.foot_words + (size_t)s * c.scene_words;

    // Prologue: the robot's pose constants (written by k_footprints) and the static tables of its type arrive by bulk copies
    // (one thread issues them, the TMA unit moves the data) while the CTA initialises its ray table and gathers footprints.
    if (tid == 0) {
        mbar_init(&sh->bar[0], 1); mbar_init(&sh->bar[1], 1); mbar_init_fence();
        mbar_expect_tx(&sh->bar[0], (unsigned)sizeof(ViewConst));
        bulk_g2s(&sh->k, d.vconst + idx, (unsigned)sizeof(ViewConst), &sh->bar[0]);
        const unsigned nb_r = ((unsigned)c.range_total * 4u + 15u) & ~15u, nb_s = ((unsigned)vh * 8u + 15u) & ~15u, nb_n = ((unsigned)c.ns * 2u + 15u) & ~15u;
        mbar_expect_tx(&sh->bar[1], nb_r + nb_s + nb_n);
        bulk_g2s(rend, d.ray_end + 2 * (size_t)ty.ray_off, nb_r, &sh->bar[1]);          // ray end cells (agent.cpp:414-430)
        bulk_g2s(spans, d.fov_spans + (size_t)ty.span_off, nb_s, &sh->bar[1]);          // FOV column spans per view row
        bulk_g2s(need, d.need_idx, nb_n, &sh->bar[1]);                                  // source rows / columns the resize reads
        sh->coll_key = 0;
        sh->red[0] = 0; sh->red[1] = 0; sh->red[2] = 0; sh->red[3] = 0; sh->red[4] = 0; sh->red[5] = 0; sh->red[6] = 0; sh->stat[0] = 0; sh->stat[1] = 0; sh->stat[2] = 0;
        sh->near_pack = 0ull; sh->n_cnear = 0; sh->n_dirty = 0; sh->n_seglist = 0;
    }
    for (int k = tid; k < c.range_total; k += VIEW_THREADS) hitkey[k] = NOHIT;
    if (tid < 64) { sh->hmin[tid] = 1023; sh->hmax[tid] = 0; }
    __syncthreads();
    mbar_wait(&sh->bar[0], 0);
    VIEW_MARK(0);      // prologue
    const bool frozen = DEBUG_FULL ? false : sh->k.frozen != 0;

    if (!frozen) {
        const unsigned H = c.H, W = c.W, Wb = c.Wb;
        const int ox = ty.org_x, oy = ty.org_y;
        const uint32_t* kpack = d.kpack + (size_t)ty.khi_off;
        const bool use_inverse = !FWD;
        const bool use_laser = FWD ? c.use_laser != 0 : true;
        const int X0 = sh->k.wbb[0], X1 = sh->k.wbb[1], Y0 = sh->k.wbb[2], Y1 = sh->k.wbb[3];
        const float i00 = (float)sh->k.inv[0], i01 = (float)sh->k.inv[1], i10 = (float)sh->k.inv[2], i11 = (float)sh->k.inv[3];
        const int orgi0 = (int)floor(sh->k.org[0]), orgi1 = (int)floor(sh->k.org[1]);
        const float orgf0 = (float)(sh->k.org[0] - orgi0), orgf1 = (float)(sh->k.org[1] - orgi1);

        // ---- Gather: footprint records of the scene whose box meets (a) the world bounding box of the field of view
        //      -> `near` (raster / ray candidates), (b) the robot's own footprint box -> `chdr` (collision candidates).
        //      A part that comes close to the FOV edge (or to the laser origin) contributes ALL of its cells, the others
        //      only their candidate cells (see phase B).
        {
            const int4 own = sh->k.own_hdr;
            const uint32_t* etiles = d.edge_tiles + ty.etile_off;
            const int etw = (vw + (1 << ET_SHIFT) - 1) >> ET_SHIFT, eth = (vh + (1 << ET_SHIFT) - 1) >> ET_SHIFT;
            int4 h_next = tid < c.NP ? __ldg(fhdr + tid) : make_int4(0, 0, 0, 0);
            for (int q = tid; q < c.NP; q += VIEW_THREADS) {
                const int4 h = h_next;
                if (q + VIEW_THREADS < c.NP) h_next = __ldg(fhdr + q + VIEW_THREADS);      // (next header in flight while this one is tested)
                if (q == r) continue;                                   // the robot never sees itself (img_env.cpp:624-628)
                const int nrow = foot_nrow(h);
                if (nrow == 0) continue;
                const int bx0 = h.x, bx1 = h.x + nrow - 1, by0 = h.y, by1 = h.y + nrow - 1;
                if (!DEBUG_FULL && bx0 <= own.x + own.z - 1 && bx1 >= own.x && by0 <= own.y + own.z - 1 && by1 >= own.y) {
                    const int p = atomicAdd(&sh->n_cnear, 1);
                    if (p < CN_CAP) { chdr[p] = h; coff[p] = d.part_off[q]; }
                }
                if (bx0 > X1 || bx1 < X0 || by0 > Y1 || by1 < Y0) continue;
                unsigned all = (!use_inverse) ? NEAR_ALL : 0u;
                {   // view-space bounding box of the part's box (+3 px): outside the FOV's pixel box -> the part cannot be seen;
                    // does it touch a tile that holds FOV-edge pixels?
                    float imin = 1e9f, imax = -1e9f, jmin = 1e9f, jmax = -1e9f;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float u = (float)(((k & 1) ? bx1 : bx0) - orgi0) - orgf0, v = (float)(((k & 2) ? by1 : by0) - orgi1) - orgf1;
                        const float qi = i00 * u + i01 * v, qj = i10 * u + i11 * v;
                        imin = fminf(imin, qi); imax = fmaxf(imax, qi); jmin = fminf(jmin, qj); jmax = fmaxf(jmax, qj);
                    }
                    if (imax + 3.f < (float)ty.fov_r0 || imin - 3.f > (float)ty.fov_r1 || jmax + 3.f < (float)ty.fov_c0 || jmin - 3.f > (float)ty.fov_c1) continue;
                  if (!all) {
                    const int ti0 = max((int)floorf(imin - 3.f) >> ET_SHIFT, 0), ti1 = min((int)floorf(imax + 3.f) >> ET_SHIFT, eth - 1);
                    const int tj0 = max((int)floorf(jmin - 3.f) >> ET_SHIFT, 0), tj1 = min((int)floorf(jmax + 3.f) >> ET_SHIFT, etw - 1);
                    if (ti1 - ti0 >= 8) all = NEAR_ALL;        // huge part: do not bother
                    else if (ti0 <= ti1 && tj0 <= tj1) {
                        // one 64-bit column mask per tile row (independent loads, no early exit)
                        const unsigned long long cols = (~0ull >> (63 - tj1)) & (~0ull << tj0);
                        unsigned long long acc = 0ull;
                        for (int ti = ti0; ti <= ti1; ti++) {
                            const uint2 row = __ldg(reinterpret_cast<const uint2*>(etiles) + ti);
                            acc |= ((unsigned long long)row.y << 32 | row.x) & cols;
                        }
                        if (acc) all = NEAR_ALL;
                    }
                  }
                }
                {   // slot in the near list and offset of the part's words among the work items of phase B, from one atomic
                    const unsigned long long t = atomicAdd(&sh->near_pack, (1ull << 32) | (unsigned long long)(nrow * foot_wpr(h)));
                    const int slot = (int)(t >> 32);
                    near[slot] = (unsigned short)(q | all); npre[slot] = (unsigned)t;
                    if (slot < NEAR_CACHE) {      // phase B then needs no dependent global loads to find the part's words
                        const int po = __ldg(d.part_off + q);
                        nhdr[slot] = h; noff[slot] = all ? po : po + ((__ldg(d.part_off + q + 1) - po) >> 1);
                    }
                }
            }
            if (use_inverse) {      // 32x32-cell world blocks under the FOV that hold static candidates
                const uint32_t* crow = d.static_crow;
                const uint32_t* orow = d.static_orow;
                const int nbj = sh->k.blk[3], nb = sh->k.blk[2] * nbj;
                for (int t = tid; t < nb; t += VIEW_THREADS) {
                    const int bi = sh->k.blk[0] + t / nbj, bj = sh->k.blk[1] + t % nbj;
                    if (__ldg(crow + (unsigned)bi * c.Wb + bj)) blocks[atomicAdd(&sh->red[3], 1)] = ((unsigned)bi << 16) | (unsigned)bj;
                    if (__ldg(orow + (unsigned)bi * c.Wb + bj)) sh->red[5] = 1;      // the static map is not empty under the FOV
                }
                if (tid >= VIEW_THREADS - 32) {     // ... nor under the robot's own footprint box (+1 cell): every block the box touches
                    const int bi0 = min(max(own.x - 1, 0), (int)H - 1) >> 5, bi1 = min(max(own.x + own.z, 0), (int)H - 1) >> 5;
                    const int bj0 = min(max(own.y - 1, 0), (int)W - 1) >> 5, bj1 = min(max(own.y + own.z, 0), (int)W - 1) >> 5;
                    const int nbw = bj1 - bj0 + 1, nbb = (bi1 - bi0 + 1) * nbw;
                    for (int u = tid - (VIEW_THREADS - 32); u < nbb; u += 32)
                        if (__ldg(orow + (unsigned)(bi0 + u / nbw) * c.Wb + (bj0 + u % nbw))) sh->red[6] = 1;
                }
            } else if (tid == 0) { sh->red[5] = 1; sh->red[6] = 1; }
        }
        __syncthreads();
        mbar_wait(&sh->bar[1], 0);
        VIEW_MARK(1);      // gather
        const int n_near = (int)(sh->near_pack >> 32);
        const unsigned n_near_words = (unsigned)sh->near_pack;
        const int n_cnear = sh->n_cnear;
        // ---- Phase A: collision code = code of the LAST colliding lattice point (agent.cpp:294-326).
        // No footprint record overlaps the robot's own box and the static map is free under it -> no point can collide.
        int best = 0;
        const bool need_A = !DEBUG_FULL && (n_cnear > 0 || sh->red[6] != 0);
        const double* pts = d.lattice_xy + 2 * (size_t)ty.pts_off;
        auto lattice_point = [&](int k) {
            double wx, wy;
            const double2 pt = __ldg(reinterpret_cast<const double2*>(pts) + k);
            tf_apply(sh->k.base_world, pt.x, pt.y, wx, wy);
            const int cx = world2cell_fast(wx, c.res, c.inv_res), cy = world2cell_fast(wy, c.res, c.inv_res);
            if ((unsigned)cx < H && (unsigned)cy < W) {
                unsigned f = 0;
                if (n_cnear <= CN_CAP) {
                    for (int e = 0; e < n_cnear; e++) { const int4 h = chdr[e]; if (foot_covers(h, fwords + coff[e], cx, cy)) f |= foot_flag(foot_kind(h)); }
                } else {     // (more colliding parts than the list holds: scan every record of the scene)
                    for (int q = 0; q < c.NP; q++) {
                        if (q == r) continue;
                        const int4 h = __ldg(fhdr + q);
                        if (foot_nrow(h) && foot_covers(h, fwords + d.part_off[q], cx, cy)) f |= foot_flag(foot_kind(h));
                    }
                }
                const int v = composed_value(__ldg(d.grid + (size_t)cx * W + cy), f);
                if (v <= 2) best = max(best, ((k + 1) << 2) | (v + 1));
            }
        };
        if (FWD && need_A) for (int k = tid; k < ty.n_pts; k += VIEW_THREADS) lattice_point(k);
        // ---- Phase B: egocentric occupancy raster (agent.cpp:373-404), 1 bit per view cell:
        // set <=> in FOV && in map && the robot's global_map_ value < 250.  FOV = static column spans per row; the pixel ->
        // world cell map is affine and evaluated in 2^-32-cell fixed point with an exact fp64 fallback inside a guard
        // band around the rounding boundary.
        //
        // (1) World -> view ("inverse") rasterisation, used when lasers are on: only raster cells that can be the first hit
        //     of a ray matter, and every such cell has a free 8-neighbour in the view (its predecessor on the ray), hence
        //     its world cell has a free cell within its 5x5 neighbourhood (a view step moves at most 2 world cells) or
        //     lies within 2 cells of the map border -- it is a CANDIDATE cell of the static map (static_cand, computed
        //     once at create) or of a footprint record (cand words) -- or the view cell sits on the FOV edge / is the laser
        //     origin (no in-FOV predecessor).  So: walk the static candidate words under the FOV (32x32 blocks with a
        //     non-zero row mask) and the words of the near footprint records, map every candidate cell back to its <= 4
        //     candidate view pixels and keep those whose EXACT forward map returns that cell.  FOV-edge pixels (static
        //     list) are evaluated forward against the static map, and parts close to FOV-edge pixels contribute all of
        //     their cells instead of their candidates.  Every bit set is a truly occupied view cell and the first hit of
        //     every ray is among them.
        // (2) Otherwise (lasers off: the "known" plane needs every FOV pixel; or a huge FOV): forward rasterisation of the
        //     static map in 32x32-pixel tiles (skipping tiles whose world footprint only touches empty blocks), then the
        //     near footprint records (all of their cells) through the same inverse mapping.
        const uint32_t* static_occ = d.static_occ;
        int* n_list2 = &sh->red[1];
        // every ray of [k0, k0+kstep, ...] within the cell's static ray interval that really passes through it
        // keeps the minimum step: hitkey[k] = min(step << 22 | cell)
        // When the rays through the cell are exactly the interval [kl, kh] (static flag, true for virtually every cell) no
        // touch test is needed and the step index is the cell's Chebyshev distance to the origin on every one of them.
        auto cell_rays = [&](unsigned cell, unsigned kp, int k0, int kstep) {      // cell = row << 16 | col
            const int pr = (int)(cell >> 16), pc = (int)(cell & 0xFFFFu);
            const int kh = kp & 0xFFFF, kl = (kp >> 16) & 0x7FFF;
            if (kp >> 31) {
                const unsigned key = hit_key(max(abs(pr - ox), abs(pc - oy)), pr, pc);
                for (int k = kl + k0; k <= kh; k += kstep) atomicMin(&hitkey[k], key);
                return;
            }
            for (int k = kl + k0; k <= kh; k += kstep) {
                const int i = ray_touch(ox, oy, rend[2 * k], rend[2 * k + 1], pr, pc);
                if (i >= 0) atomicMin(&hitkey[k], hit_key(i, pr, pc));
            }
        };
        // An occupied raster cell updates the rays through it on the spot (phase C); only cells crossed by many rays (close
        // to the origin) are listed and shared out over a warp after the barrier.
        auto push_cell_kp = [&](int pr, int pc, unsigned kp) {           // kp = the cell's ray interval (kpack entry)
            const int kh = kp & 0xFFFF, kl = (kp >> 16) & 0x7FFF;
            if (kh == 0xFFFF) return;                                      // no ray passes through this cell
            if (VIEW_STATS) atomicAdd(&sh->stat[1], 1);
            if (kh - kl + 1 > BL_HEAVY) {
                const int p = atomicAdd(n_list2, 1);
                if (p < BL2_CAP) { blist2[p] = ((unsigned)pr << 16) | (unsigned)pc; return; }
            }
            if (kp >> 31) {
                const unsigned key = hit_key(max(abs(pr - ox), abs(pc - oy)), pr, pc);
                for (int k = kl; k <= kh; k++) atomicMin(&hitkey[k], key);
            } else cell_rays_inline(hitkey, rend, ox, oy, pr, pc, kl, kh);
        };
        auto push_cell = [&](int pr, int pc) { push_cell_kp(pr, pc, __ldg(kpack + pr * vw + pc)); };
        // The same one cell late: the table load of a cell is in flight while the thread works out its next candidate.
        unsigned pend_cell = 0xFFFFFFFFu, pend_kp = 0u;
        auto push_cell_deferred = [&](int pr, int pc) {
            const unsigned kp = __ldg(kpack + pr * vw + pc);
            if (pend_cell != 0xFFFFFFFFu) push_cell_kp((int)(pend_cell >> 16), (int)(pend_cell & 0xFFFFu), pend_kp);
            pend_cell = ((unsigned)pr << 16) | (unsigned)pc; pend_kp = kp;
        };
        const int n_trow = (vh + 31) >> 5, n_tiles = n_trow * vwb;
        if (!use_inverse) for (int q = tid; q < vh * vwb; q += VIEW_THREADS) { occ[q] = 0u; if (!use_laser) known[q] = 0u; }
        if (!use_inverse) {
            // forward rasterisation of the static map
            const uint32_t* orow = d.static_orow;
            int* n_active = &sh->red[2];
            unsigned short* tile_list = reinterpret_cast<unsigned short*>(blist2);     // region B is free until the heavy-cell list fills
            for (int t = tid; t < n_tiles; t += VIEW_THREADS) {
                if (!((d.tile_fov[ty.tile_off + (t >> 5)] >> (t & 31)) & 1u)) continue;
                bool active = !use_laser;          // the "known" plane needs every FOV pixel
                if (!active) {
                    const int ti = t / vwb, tj = t - ti * vwb;
                    const int i0 = ti * 32, i1 = min(i0 + 31, vh - 1), j0 = tj * 32, j1 = min(j0 + 31, vw - 1);
                    int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int ii = (k & 1) ? i1 : i0, jj = (k & 2) ? j1 : j0;
                        const int cx = (int)((sh->k.cx + (long long)ii * sh->k.ax + (long long)jj * sh->k.bx) >> 32);
                        const int cy = (int)((sh->k.cy + (long long)ii * sh->k.ay + (long long)jj * sh->k.by) >> 32);
                        xmin = min(xmin, cx); xmax = max(xmax, cx); ymin = min(ymin, cy); ymax = max(ymax, cy);
                    }
                    xmin = max(xmin - 1, 0); ymin = max(ymin - 1, 0); xmax = min(xmax + 1, (int)H - 1); ymax = min(ymax + 1, (int)W - 1);
                    for (int bi = xmin >> 5; bi <= (xmax >> 5) && !active; bi++)
                        for (int bj = ymin >> 5; bj <= (ymax >> 5); bj++)
                            if (__ldg(orow + (unsigned)bi * Wb + bj)) { active = true; break; }
                }
                if (active) { const int ti = t / vwb; tile_list[atomicAdd(n_active, 1)] = (unsigned short)((ti << 8) | (t - ti * vwb)); }
            }
        }
        if (!use_inverse) __syncthreads();       // (world->view mode: block list and near list were finished before the last barrier)
        if (!use_inverse) {
            const int n_items = sh->red[2] * 32;
            const long long lbx = sh->k.cx + (long long)lane * sh->k.bx, lby = sh->k.cy + (long long)lane * sh->k.by;
            const long long bx32 = sh->k.bx * 32, by32 = sh->k.by * 32;
#pragma unroll 2
            for (int item = warp; item < n_items; item += VIEW_THREADS / 32) {
                const int t = reinterpret_cast<unsigned short*>(blist2)[item >> 5];
                const int wj = t & 255;
                const int i = (t >> 8) * 32 + (item & 31);
                if (i >= vh) continue;
                const int a0 = spans[i * 4 + 0], a1 = spans[i * 4 + 1], b0 = spans[i * 4 + 2], b1 = spans[i * 4 + 3];
                const int j = wj * 32 + lane;
                const bool in_fov = (j >= a0 && j < a1) || (j >= b0 && j < b1);     // empty spans are (-1,-1)
                bool o = false, kn = false;
                if (in_fov) {
                    const long long tx = lbx + (long long)i * sh->k.ax + (long long)wj * bx32;
                    const long long tyy = lby + (long long)i * sh->k.ay + (long long)wj * by32;
                    int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
                    const unsigned lx = (unsigned)tx, ly = (unsigned)tyy;
                    if (lx + FX_GUARD < 2 * FX_GUARD || ly + FX_GUARD < 2 * FX_GUARD) {
                        exact_cell(sh->k.view_world, c.res, i, j, cx, cy);
                    }
                    if ((unsigned)cx < H && (unsigned)cy < W) {
                        kn = true;
                        o = (__ldg(static_occ + (unsigned)cx * Wb + ((unsigned)cy >> 5)) >> (cy & 31)) & 1u;
                    }
                }
                const unsigned wo = __ballot_sync(0xffffffffu, o);
                if (!use_laser) { const unsigned wk = __ballot_sync(0xffffffffu, kn); if (lane == 0) known[i * vwb + wj] = wk; }
                if (lane == 0) occ[i * vwb + wj] = wo;
            }
            __syncthreads();      // the footprint records below OR into the words written above
        }
        // FOV-edge pixels (and the laser origin): forward, static map only -- footprint records near them come in whole below
        const uint32_t* edge = d.edge_px + ty.edge_off;
        auto edge_pixel = [&](int e) {
            const unsigned ep = __ldg(edge + e);
            const int i = ep >> 16, j = ep & 0xFFFF;
            const long long tx = sh->k.cx + (long long)i * sh->k.ax + (long long)j * sh->k.bx;
            const long long tyy = sh->k.cy + (long long)i * sh->k.ay + (long long)j * sh->k.by;
            int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
            if ((unsigned)tx + FX_GUARD < 2 * FX_GUARD || (unsigned)tyy + FX_GUARD < 2 * FX_GUARD) {
                exact_cell(sh->k.view_world, c.res, i, j, cx, cy);
            }
            if ((unsigned)cx < H && (unsigned)cy < W) {
                const bool o = (__ldg(static_occ + (unsigned)cx * Wb + ((unsigned)cy >> 5)) >> (cy & 31)) & 1u;
                if (o) push_cell(i, j);
            }
        };
        // the light, even parts of the phase (world->view mode): FOV-edge pixels when the static map is not empty under the
        // FOV, and the collision lattice
        if (use_inverse && sh->red[5]) for (int e = tid; e < ty.n_edge; e += VIEW_THREADS) edge_pixel(e);
        if (use_inverse && need_A) for (int k = tid; k < ty.n_pts; k += VIEW_THREADS) lattice_point(k);
        {
            // candidate words -> candidate cells -> view pixels.  Work items: 32 rows per listed static block, then every
            // word of every near footprint record.  Dense words make the work per word very uneven, so the CTA takes the
            // items in chunks of two per thread: (1) every thread decodes and loads its two words into shared memory, (2) a
            // block-wide prefix sum numbers the candidate bits of the chunk, (3) the candidates are split EVENLY over the
            // threads -- each thread finds the word holding its first candidate by bisection and then walks on bit by bit.
            const uint32_t* static_cand = d.static_cand;
            const int n_static = use_inverse ? sh->red[3] * 32 : 0;
            const int n_items = n_static + (int)n_near_words;
            const float f00 = (float)sh->k.view_world.m00, f01 = (float)sh->k.view_world.m01, f10 = (float)sh->k.view_world.m10, f11 = (float)sh->k.view_world.m11;
            const float fr0 = (float)ty.fov_r0 - 1.5f, fr1 = (float)ty.fov_r1 + 1.5f, fc0 = (float)ty.fov_c0 - 1.5f, fc1 = (float)ty.fov_c1 + 1.5f;
            auto candidate = [&](int cX, int cY) {
                const float u = (float)(cX - orgi0) - orgf0, v = (float)(cY - orgi1) - orgf1;      // cell - org, exact integer part
                const float qi = i00 * u + i01 * v, qj = i10 * u + i11 * v;
                // Pixel (i,j) maps to this cell iff M*((i,j) - q) lies in the unit square around the cell centre (M =
                // rotation of view_world).  The float test decides all pixels farther than INV_EPS from the square's
                // edge; only the others run the exact forward map.  |q| < 2^10 so the float error is < 1e-3 cell.
                if (qi < fr0 || qi > fr1 || qj < fc0 || qj > fc1) return;       // (pixel box of the FOV, 1.5 px margin)
                const int ia = (int)ceilf(qi - 0.72f), ja = (int)ceilf(qj - 0.72f);
                // cheap part for the 2 x 2 window at once (M*d is linear: the four offsets share two products), ...
                const float du0 = (float)ia - qi, dv0 = (float)ja - qj;
                const float x00 = f00 * du0 + f01 * dv0, y00 = f10 * du0 + f11 * dv0;
                float e4[4];
                e4[0] = fmaxf(fabsf(x00), fabsf(y00));
                e4[1] = fmaxf(fabsf(x00 + f01), fabsf(y00 + f11));
                e4[2] = fmaxf(fabsf(x00 + f00), fabsf(y00 + f10));
                e4[3] = fmaxf(fabsf(x00 + f00 + f01), fabsf(y00 + f10 + f11));
                unsigned live = (e4[0] <= 0.5f + INV_EPS ? 1u : 0u) | (e4[1] <= 0.5f + INV_EPS ? 2u : 0u) |
                                (e4[2] <= 0.5f + INV_EPS ? 4u : 0u) | (e4[3] <= 0.5f + INV_EPS ? 8u : 0u);
                // ... then the (usually one) surviving pixel
                while (live) {
                    const int t = __ffs(live) - 1; live &= live - 1;
                    const int i = ia + (t >> 1), j = ja + (t & 1);
                    const float em = t == 0 ? e4[0] : t == 1 ? e4[1] : t == 2 ? e4[2] : e4[3];
                    if ((unsigned)i >= (unsigned)vh || (unsigned)j >= (unsigned)vw) continue;
                    const int a0 = spans[i * 4 + 0], a1 = spans[i * 4 + 1], b0 = spans[i * 4 + 2], b1 = spans[i * 4 + 3];
                    if (!((j >= a0 && j < a1) || (j >= b0 && j < b1))) continue;
                    if (em > 0.5f - INV_EPS) {
                        const long long tx = sh->k.cx + (long long)i * sh->k.ax + (long long)j * sh->k.bx;
                        const long long tyy = sh->k.cy + (long long)i * sh->k.ay + (long long)j * sh->k.by;
                        int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
                        if ((unsigned)tx + FX_GUARD < 2 * FX_GUARD || (unsigned)tyy + FX_GUARD < 2 * FX_GUARD) {
                            exact_cell(sh->k.view_world, c.res, i, j, cx, cy);
                        }
                        if (cx != cX || cy != cY) continue;
                    }
                    // World->view mode: no raster is kept -- a view pixel has ONE world cell, so it can only be found twice when
                    // two records (or a record and the static map) cover that cell; it is then listed twice, which is harmless.
                    // Forward mode: into the raster; its boundary cells are listed by the scan below.
                    if (use_inverse) push_cell_deferred(i, j);
                    else atomicOr(&occ[i * vwb + (j >> 5)], 1u << (j & 31));
                }
            };
            for (int c0 = 0; c0 < n_items; c0 += CAND_CHUNK) {
                const int n_ch = min(CAND_CHUNK, n_items - c0);
                int cnt2[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int it = 2 * tid + u, item = c0 + it;
                    unsigned cand = 0; int X = 0, bj = 0;
                    if (it < n_ch) {
                        if (item < n_static) {
                            const unsigned bb = blocks[item >> 5];
                            X = (int)(bb >> 16) * 32 + (item & 31); bj = (int)(bb & 0xFFFF);
                            if (X >= X0 && X <= X1) cand = __ldg(static_cand + (unsigned)X * Wb + bj);
                        } else {
                            const unsigned di = (unsigned)(item - n_static);
                            int lo = 0, hi = n_near - 1;                       // last k with npre[k] <= di
                            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (npre[mid] <= di) lo = mid; else hi = mid - 1; }
                            int4 h; int wbase;
                            if (lo < NEAR_CACHE) { h = nhdr[lo]; wbase = noff[lo]; }
                            else {
                                const unsigned e = near[lo];
                                const int q = e & 0x7FFF;
                                h = __ldg(fhdr + q);
                                const int po = __ldg(d.part_off + q);
                                wbase = (e & NEAR_ALL) ? po : po + ((__ldg(d.part_off + q + 1) - po) >> 1);
                            }
                            const int wpr = foot_wpr(h), wi = (int)(di - npre[lo]);
                            const int rr = wi / wpr;
                            X = h.x + rr; bj = foot_wj0(h) + (wi - rr * wpr);
                            if (X >= X0 && X <= X1) cand = __ldg(fwords + wbase + wi);
                        }
                        if (cand) {     // only columns under the FOV's bounding box
                            const int lo = max(Y0 - bj * 32, 0), hi = min(Y1 - bj * 32, 31);
                            cand = lo <= hi ? cand & ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo)) : 0u;
                        }
                    }
                    cword[it] = cand; cmeta[it] = ((unsigned)X << 12) | (unsigned)bj;
                    cnt2[u] = __popc(cand);
                }
                const int mine = cnt2[0] + cnt2[1];
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (lane == 31) cwsum[warp] = incl;
                __syncthreads();
                int woff = 0, total = 0;
#pragma unroll
                for (int w = 0; w < VIEW_THREADS / 32; w++) { const int v = cwsum[w]; total += v; if (w < warp) woff += v; }
                const int excl = woff + incl - mine;
                cpre[2 * tid] = (unsigned short)excl; cpre[2 * tid + 1] = (unsigned short)(excl + cnt2[0]);
                __syncthreads();
                if (VIEW_STATS && tid == 0) atomicAdd(&sh->stat[0], total);
                const int per = (total + VIEW_THREADS - 1) / VIEW_THREADS;
                int k = tid * per;
                const int k_end = min(k + per, total);
                if (k < k_end) {
                    int lo = 0, hi = n_ch - 1;                                 // last word with cpre[w] <= k (empty words share their successor's count)
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)cpre[mid] <= k) lo = mid; else hi = mid - 1; }
                    int wi = lo;
                    unsigned m = cword[wi];
                    {   // drop the candidates of this word that belong to the thread before
                        const int skip = k - (int)cpre[wi];
                        if (skip) m &= 0xFFFFFFFEu << nth_set_bit(m, skip - 1);
                    }
                    unsigned meta = cmeta[wi];
                    for (; k < k_end; k++) {
                        while (m == 0u) { wi++; m = cword[wi]; meta = cmeta[wi]; }
                        const int bit = __ffs(m) - 1; m &= m - 1;
                        candidate((int)(meta >> 12), (int)(meta & 0xFFFu) * 32 + bit);
                    }
                }
                if (c0 + CAND_CHUNK < n_items) __syncthreads();       // the chunk arrays are rewritten
            }
            if (pend_cell != 0xFFFFFFFFu) push_cell_kp((int)(pend_cell >> 16), (int)(pend_cell & 0xFFFFu), pend_kp);
        }
        if (!DEBUG_FULL && need_A) {
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (lane == 0 && best) atomicMax(&sh->coll_key, best);
        }
        __syncthreads();
        VIEW_MARK(2);      // phase B (+ A, edge pixels)
        if (!DEBUG_FULL && tid == 0) {
            int code = sh->coll_key & 3;
            RBF(d, RB_COLL, idx) = (double)code;
            view_publish_code(d, idx, code, is_reset);
        }

        // ---- Phase C: first occupied cell of every laser ray (agent.cpp:405-438, 511-624).
        // A ray's hit cell always has a free 8-neighbour (its predecessor on the ray), so only the boundary
        // cells of the raster can be hits.  For each boundary cell the (static) interval of ray indices whose
        // integer line walk passes through it is scanned with the closed-form touch test and the ray keeps
        // the minimum step (atomicMin on step<<22|cell).  Cells that do not fit the lists are resolved inline.
        int any_hit_all = 0;
        if (use_laser) {
            if (!use_inverse) {
                for (int q = tid; q < vh * 16 * ((vwb + 15) / 16); q += VIEW_THREADS) {
                    const int wpr = 16 * ((vwb + 15) / 16);                  // words per row rounded up to 16: shift/mask indexing
                    const int i = (wpr == 16) ? (q >> 4) : q / wpr, wj = (wpr == 16) ? (q & 15) : q - i * wpr;
                    if (wj >= vwb) continue;
                    const unsigned O = occ[i * vwb + wj];
                    if (!O) continue;
                    unsigned all8 = 0xffffffffu;
                    for (int di = -1; di <= 1; di++) {
                        const int ii = i + di;
                        unsigned m = 0, ml = 0, mr = 0;      // outside the raster counts as free (conservative superset)
                        if (ii >= 0 && ii < vh) {
                            m = occ[ii * vwb + wj];
                            ml = wj > 0 ? occ[ii * vwb + wj - 1] : 0u;
                            mr = wj + 1 < vwb ? occ[ii * vwb + wj + 1] : 0u;
                        }
                        const unsigned left = (m << 1) | (ml >> 31), right = (m >> 1) | (mr << 31);
                        all8 &= left & right;
                        if (di != 0) all8 &= m;
                    }
                    unsigned bnd = O & ~all8;
                    if (i == ox && (oy >> 5) == wj) bnd |= O & (1u << (oy & 31));   // an occupied origin hits every ray at step 0
                    while (bnd) {
                        const int b = __ffs(bnd) - 1; bnd &= bnd - 1;
                        const int col = wj * 32 + b;
                        if (col >= vw) break;
                        push_cell(i, col);
                    }
                }
                __syncthreads();
            }
            const int nl2 = min(sh->red[1], BL2_CAP);
            if (d.dbg_stats && tid == 0) {
                int* st = d.dbg_stats + 4 * (size_t)idx;
                st[0] = sh->red[2] + sh->red[3] + n_near; st[1] = sh->red[0]; st[2] = sh->red[1];
                st[3] = sh->red[1] > BL2_CAP;
            }
            for (int q = warp; q < nl2; q += VIEW_THREADS / 32) { const unsigned cell = blist2[q]; cell_rays(cell, __ldg(kpack + (cell >> 16) * vw + (cell & 0xFFFFu)), lane, 32); }
            __syncthreads();
            if (!DEBUG_FULL) {
                // laser ranges out; one bit per ray "hit something" (ballots: a warp's 32 rays are one word) and, per block of
                // rays, the nearest hit -- the output classification below asks both about ray intervals
                int any_local = 0;
                const float nohit_out = c.laser_norm ? (float)(6.0 / c.laser_max) : 6.f;     // hit = 6 without a hit (agent.cpp:513)
                for (int k0 = 0; k0 < c.range_total; k0 += VIEW_THREADS) {
                    const int k = k0 + tid;
                    const bool valid = k < c.range_total;
                    const unsigned key = valid ? hitkey[k] : NOHIT;
                    const bool hit_any = key != NOHIT;
                    if (valid) {
                        float out = nohit_out;
                        if (hit_any) {
                            const int hx = (key >> 11) & 2047, hy = key & 2047;
                            double x0 = ox * c.res, y0 = oy * c.res, xc = hx * c.res, yc = hy * c.res;
                            const float wire = (float)sqrt((x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc));   // AgentState.laser is float32[]
                            out = c.laser_norm ? (float)((double)wire / c.laser_max) : wire;            // yaml_env.py:440-444
                        }
                        d.o_laser[(size_t)idx * c.range_total + k] = out;
                    }
                    const unsigned word = __ballot_sync(0xffffffffu, hit_any);
                    if (lane == 0 && (k >> 5) <= ((c.range_total - 1) >> 5)) hbits[k >> 5] = word;
                    any_local |= word != 0u;
                    if (c.hb_shift == 4) {      // blocks of 16 rays = half warps: shuffle minimum / maximum, no atomics
                        // redux.sync with the full (uniform) mask, the other half warp neutralised -- a per-lane half mask makes
                        // the compiler fall back to a loop
                        const bool lo_half = lane < 16;
                        const int hp0 = (int)(key >> 22), hq0 = valid ? hp0 : 0;
                        const int mn_lo = __reduce_min_sync(0xffffffffu, lo_half ? hp0 : 0x7fffffff), mn_hi = __reduce_min_sync(0xffffffffu, lo_half ? 0x7fffffff : hp0);
                        const int mx_lo = __reduce_max_sync(0xffffffffu, lo_half ? hq0 : 0), mx_hi = __reduce_max_sync(0xffffffffu, lo_half ? 0 : hq0);
                        if ((lane & 15) == 0 && valid) { sh->hmin[k >> 4] = lo_half ? mn_lo : mn_hi; sh->hmax[k >> 4] = lo_half ? mx_lo : mx_hi; }
                    } else if (valid) {
                        if (hit_any) atomicMin(&sh->hmin[k >> c.hb_shift], (int)(key >> 22));
                        atomicMax(&sh->hmax[k >> c.hb_shift], (int)(key >> 22));
                    }
                }
                any_hit_all = __syncthreads_or(any_local);
                VIEW_MARK(3);      // heavy cells + laser ranges
            }
        }

        // ---- Phase D/E/F: laser_map reconstruction fused into the cubic resize, per OUTPUT pixel.
        // D/E: the final view_map_ value of a pixel = last-writer-wins over the rays in index order, evaluated from
        //      the highest touching ray downwards (static tables), then the robot's own footprint (100, agent.cpp:503).
        //      dtab packs, per pixel the resize reads, the top ray, its step index there and the own-footprint bit:
        //      the top ray touches by construction, so the closed-form touch test only runs on fall-through.
        // F:   cv2.resize(INTER_CUBIC) 400->48 (yaml_env.py:433-434), OpenCV's own path: horizontal pass in int32 with
        //      11-bit weights, vertical pass as an fp32 FMA chain with weights * 2^-22, round-half-even, saturate; then
        //      float16(x)/255 via a host-built table.  The scale is 8.33 > 4 taps, so the <= 16 source pixels of an output
        //      pixel belong to it alone: an output whose source pixels' top rays ALL missed (prefix counts over the static
        //      ray interval okk, O(1)) keeps its hit-free value, a table entry (oval).  Only the other ("dirty") outputs are
        //      evaluated: 4 threads per output, one per source row, combined with shuffles.
        const uint32_t* own_mask = d.own_mask + (size_t)ty.own_mask_off;
        const uint32_t* dtab = d.dtab + (size_t)ty.dtab_off;
        // value code of tap k of output column oc on needed row rr: 0 -> 0, 1 -> 100, 2 -> 200, 3 -> 255.
        // e is the tap's dtab entry; the source column is only looked up on the rare fall-through path.
        auto pixel_code = [&](unsigned e, int rr, const short* tp, int k) -> unsigned {
            unsigned code = 2u; bool own;
            if (use_laser) {
                own = e >> 31;
                const int kh = e & 0xFFF;
                if (kh != 0xFFF) {
                    const unsigned key = hitkey[kh];
                    const int hp = (int)(key >> 22), i0 = (e >> 12) & 0x3FF;
                    if (i0 < hp) code = 3u;
                    else if (i0 == hp) code = 0u;
                    else {
                        const int pr = need[rr], pc = need[tp[k]];
                        if (!(pr != (int)((key >> 11) & 2047) && pc != (int)(key & 2047))) {     // no shadow write: fall through to lower rays
                            const unsigned kp = __ldg(kpack + pr * vw + pc);
                            code = pixel_code_below(hitkey, rend, ox, oy, pr, pc, kh, (int)((kp >> 16) & 0x7FFF));
                        }
                    }
                }
            } else {
                const int pr = need[rr], pc = need[tp[k]], full = pr * vw + pc;
                const bool o = (occ[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                const bool kn = (known[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                code = o ? 0u : (kn ? 3u : 2u);
                own = (own_mask[full >> 5] >> (full & 31)) & 1u;
            }
            if (code != 0u && own) code = 1u;
            return code;
        };
        if (DEBUG_FULL) {
            // whole 400x400 raster for the tests (generic per-pixel path)
            for (int rr = warp; rr < vh; rr += VIEW_THREADS / 32) {
                for (int cc = lane; cc < vw; cc += 32) {
                    const int pr = rr, pc = cc, full = pr * vw + pc;
                    int val = 200;
                    if (use_laser) {
                        const unsigned kp = __ldg(kpack + full);
                        const int kh = kp & 0xFFFF;
                        if (kh != 0xFFFF) {
                            const int kl = (kp >> 16) & 0x7FFF;
                            for (int k = kh; k >= kl; k--) {
                                const int i = ray_touch(ox, oy, rend[2 * k], rend[2 * k + 1], pr, pc);
                                if (i < 0) continue;
                                const unsigned key = hitkey[k];
                                const int hp = (int)(key >> 22);
                                if (i < hp) { val = 255; break; }
                                if (i == hp) { val = 0; break; }
                                if (pr != (int)((key >> 11) & 2047) && pc != (int)(key & 2047)) { val = 200; break; }   // shadow write (agent.cpp:557-558)
                            }
                        }
                    } else {
                        const bool o = (occ[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                        const bool kn = (known[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                        val = o ? 0 : (kn ? 255 : 200);
                    }
                    if (val != 0 && pr >= ty.zone_r0 && pr <= ty.zone_r1 && pc >= ty.zone_c0 && pc <= ty.zone_c1 &&
                        ((own_mask[full >> 5] >> (full & 31)) & 1u)) val = 100;
                    if (d.dbg_view) d.dbg_view[(size_t)idx * vh * vw + full] = (uint8_t)val;
                }
            }
        } else {
            uint16_t* o_img = d.o_sensor + (size_t)idx * npx;
            const uint32_t* oshad = d.ostat + (size_t)ty.ostat_off + npx + (npx + 1) / 2;               // [npx] all-shadow test: ray interval | nearest source pixel
            const uint16_t* oshval = reinterpret_cast<const uint16_t*>(oshad + npx);                    // [npx] all-shadow float16
            const bool any_hit = !use_laser || any_hit_all != 0;
            // any ray of [a, b] with a hit?  (a, b at most a few words apart)
            auto range_hit = [&](int a, int b) -> bool {
                const int wa = a >> 5, wb = b >> 5;
                const unsigned ma = 0xffffffffu << (a & 31), mb = 0xffffffffu >> (31 - (b & 31));
                if (wa == wb) return (hbits[wa] & ma & mb) != 0u;
                unsigned acc = (hbits[wa] & ma) | (hbits[wb] & mb);
                for (int w = wa + 1; w < wb; w++) acc |= hbits[w];
                return acc != 0u;
            };
            // (1) Segments of 8 consecutive outputs (one 16-byte store): a segment none of whose rays was stopped in front of
            //     its farthest source pixel keeps its hit-free values -- one table load, one store.  The others are listed.
            const int n_seg = (npx + 7) >> 3;
            const bool seg_ok = use_laser && (npx & 7) == 0 && (reinterpret_cast<size_t>(d.o_sensor) & 15) == 0;     // (16-byte stores)
            const uint32_t* oseg = oshad + npx + (npx + 1) / 2;        // [n_seg] like okk, over the segment's outputs
            for (int sg = tid; sg < n_seg; sg += VIEW_THREADS) {
                bool flagged = !seg_ok;
                uint4 free8 = make_uint4(0u, 0u, 0u, 0u);
                if (seg_ok) free8 = __ldg(reinterpret_cast<const uint4*>(oval) + sg);      // (in flight beside the test's own table load)
                if (seg_ok && any_hit) {
                    const unsigned kk = __ldg(oseg + sg);
                    const int kmin = kk & 0xFFFu, kmax = (kk >> 12) & 0xFFFu;
                    if (kmax >= kmin) {
                        const int b0 = kmin >> c.hb_shift, b1 = kmax >> c.hb_shift;
                        int hm = b1 - b0 < 16 ? 1023 : 0;                 // (segments next to the origin: listed without asking)
                        if (hm) for (int b = b0; b <= b1; b++) hm = min(hm, sh->hmin[b]);
                        flagged = hm <= (int)(kk >> 24) * 4;
                    }
                }
                if (!flagged) reinterpret_cast<uint4*>(o_img)[sg] = free8;
                else seglist[atomicAdd(&sh->n_seglist, 1)] = (unsigned short)sg;
            }
            __syncthreads();
            VIEW_MARK(4);      // D1 segments
            // (2) the outputs of the listed segments, one per thread
            const int n_listed = sh->n_seglist * 8;
            for (int it = tid; it < n_listed; it += VIEW_THREADS) {
                const int q = (int)seglist[it >> 3] * 8 + (it & 7);
                if (q >= npx) continue;
                bool is_dirty = !use_laser;
                // (the four table entries of the output are loaded together: one memory latency instead of three in a row)
                const unsigned kk = use_laser ? __ldg(okk + q) : 0u, ks = use_laser ? __ldg(oshad + q) : 0u;
                const uint16_t v_free = __ldg(oval + q), v_shadow = use_laser ? __ldg(oshval + q) : (uint16_t)0;
                if (use_laser) {
                    const int kmin = kk & 0xFFFu, kmax = (kk >> 12) & 0xFFFu;
                    if (kmax >= kmin && range_hit(kmin, kmax)) {
                        // some of the rays hit something: still clean if every hit lies beyond all of the output's source pixels
                        // (they are then all "free", as in the hit-free value) -- nearest hit over the covering ray blocks.
                        // (Outputs next to the origin see hundreds of rays: evaluated in full without asking.)
                        const int b0 = kmin >> c.hb_shift, b1 = kmax >> c.hb_shift;
                        is_dirty = true;
                        if (b1 - b0 < 8) {
                            int hm = 1023;
                            for (int b = b0; b <= b1; b++) hm = min(hm, sh->hmin[b]);
                            is_dirty = hm <= (int)(kk >> 24) * 4;
                        }
                        if (is_dirty) {
                            // ... or if EVERY ray through a source pixel was stopped in front of it: each source pixel is then
                            // "unknown" (200: shadow-written or never written, agent.cpp:557-558) whatever the rays' order -> the
                            // output takes its all-shadow value, another table entry
                            const int r0 = (int)(ks & 0xFFFu), r1 = (int)((ks >> 12) & 0xFFFu), a0 = r0 >> c.hb_shift, a1 = r1 >> c.hb_shift;
                            if (r1 >= r0 && a1 - a0 < 8) {
                                int hx = 0;
                                for (int b = a0; b <= a1; b++) hx = max(hx, sh->hmax[b]);
                                if (hx < (int)(ks >> 24) * 4) { if (VIEW_STATS) atomicAdd(&sh->stat[2], 1); o_img[q] = v_shadow; continue; }
                            }
                        }
                    }
                }
                if (is_dirty) dirty[atomicAdd(&sh->n_dirty, 1)] = (unsigned short)q;
                else o_img[q] = v_free;
            }
            __syncthreads();
            VIEW_MARK(5);      // D2 listed outputs
            const int n4 = sh->n_dirty * 4;
            if (VIEW_STATS && tid == 0) {
                unsigned long long* st = d.counters + 4;
                atomicAdd(st + 0, 1ull); atomicAdd(st + 1, (unsigned long long)(sh->near_pack >> 32)); atomicAdd(st + 2, (unsigned long long)(unsigned)sh->near_pack);
                atomicAdd(st + 3, (unsigned long long)sh->red[3]); atomicAdd(st + 4, (unsigned long long)sh->stat[0]); atomicAdd(st + 5, (unsigned long long)sh->stat[1]);
                atomicAdd(st + 6, (unsigned long long)sh->n_dirty); atomicAdd(st + 7, (unsigned long long)sh->stat[2]); atomicAdd(st + 8, (unsigned long long)(sh->n_cnear > 0 || sh->red[6]));
                atomicAdd(st + 9, (unsigned long long)sh->red[5]); atomicAdd(st + 10, (unsigned long long)sh->red[1]); atomicAdd(st + 11, (unsigned long long)(any_hit_all != 0));
                atomicAdd(st + 12, (unsigned long long)sh->n_seglist);
            }
            const float scale = 1.f / (2048.f * 2048.f);
            for (int base = warp * 32; base < n4; base += VIEW_THREADS) {
                const int it = base + lane;
                const bool act = it < n4;
                const int q = act ? dirty[it >> 2] : 0, t = it & 3;
                const int orow = (int)__umulhi((unsigned)q, c.img_inv), oc = q - orow * c.img;
                float sv = 0.f;
                if (act) {
                    const short* tp = ctap + 4 * oc;
                    const short4 cf = __ldg(reinterpret_cast<const short4*>(ccoef) + oc);
                    const int rr = ctap[4 * orow + t];
                    uint4 e4 = make_uint4(0u, 0u, 0u, 0u);
                    if (use_laser) e4 = __ldg(reinterpret_cast<const uint4*>(dtab) + rr * c.img + oc);
                    int acc = 0;
                    if (cf.x) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.x, rr, tp, 0))) & 0xFFu) * cf.x;
                    if (cf.y) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.y, rr, tp, 1))) & 0xFFu) * cf.y;
                    if (cf.z) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.z, rr, tp, 2))) & 0xFFu) * cf.z;
                    if (cf.w) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.w, rr, tp, 3))) & 0xFFu) * cf.w;
                    sv = (float)acc;
                }
                const float s1 = __shfl_down_sync(0xffffffffu, sv, 1), s2 = __shfl_down_sync(0xffffffffu, sv, 2), s3 = __shfl_down_sync(0xffffffffu, sv, 3);
                if (act && t == 0) {
                    const short4 cv = __ldg(reinterpret_cast<const short4*>(ccoef) + orow);
                    const float b0 = cv.x * scale, b1 = cv.y * scale, b2 = cv.z * scale, b3 = cv.w * scale;
                    const float v = fmaf(sv, b0, fmaf(s1, b1, fmaf(s2, b2, s3 * b3)));
                    int iv = __float2int_rn(v);
                    iv = min(255, max(0, iv));
                    o_img[q] = __ldg(lut16 + iv);
                }
            }
        }
    }
    VIEW_MARK(6);          // dirty outputs (thread 0's share)
    if (frozen) {
        mbar_wait(&sh->bar[1], 0);       // (a CTA must not exit with bulk copies in flight)
        if (!DEBUG_FULL && tid == 0) view_publish_code(d, idx, (int)RBF(d, RB_COLL, idx), is_reset);      // stale code re-sent (agent.cpp:358-360)
    }
    // (Phase G, the state vector, is written by k_view_consts; the pedestrian observation by k_ped_obs on its own stream.)
}

// ---------------------------------------------------------------------------------------------
// Pedestrian observation of every robot (img_env.cpp:566-583 ped_info; yaml_env.py:392-466 _get_states /
// _draw_ped_map): pedestrians in the robot frame, nearest first (stable sort), ped_vector_states, ped_min_dists
// (NearbyPed persistence) and the 3 x img x img ped_maps where farther pedestrians overwrite nearer ones.
// It only reads poses, so it runs beside the stamp / view kernels on the library's side stream.
// ---------------------------------------------------------------------------------------------
#define PED_THREADS 128
inline PedLayout ped_layout(const Cfg& c) {
    PedLayout L;
    int n = 1; while (n < c.P) n <<= 1;
    L.n_sort = n;
    size_t off = 0;
    L.winner = off; off += ((size_t)c.img * c.img * 4 + 15) & ~(size_t)15;
    L.keys = off; off += ((size_t)n * 8 + 15) & ~(size_t)15;
    L.pobs = off; off += (size_t)(c.P > 0 ? c.P : 1) * 16;
    L.row = off; off += ((size_t)c.pvs_len * 4 + 15) & ~(size_t)15;
    L.dkeys = off; off += ((size_t)n * 8 + 15) & ~(size_t)15;
    L.paint = off; off += (size_t)(c.P > 0 ? c.P : 1) * 16;
    L.total = off + 16;
    return L;
}
inline size_t ped_smem_bytes(const Cfg& c) { return ped_layout(c).total; }

__global__ void __launch_bounds__(PED_THREADS) k_ped_obs(Dev d, const int* scene_ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Cfg& c = d.c;
    const int tid = threadIdx.x;
    const int sl = blockIdx.x / c.R, r = blockIdx.x % c.R;
    if (d.n_dev && sl >= *d.n_dev) return;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int idx = s * c.R + r;
    const RobotType& ty = d.types[d.type_of[r]];
    const PedLayout& L = d.pl;
    int* winner = reinterpret_cast<int*>(smem_raw + L.winner);
    // Sort keys: python sorts by the float64 x*x + y*y, stably.  The network sorts ONE 64-bit word per pedestrian, float32(key)
    // << 32 | index (rounding to float32 is monotone, so the order can only be wrong inside a run of equal float32 keys: such
    // runs -- rare -- are re-sorted by (float64 key, index) afterwards).
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + L.keys);     // float32 bits of the key << 32 | pedestrian index
    int4* paint = reinterpret_cast<int4*>(smem_raw + L.paint);                               // pedestrians inside the ped map: rank << 16 | index, cell window
    double* dkeys = reinterpret_cast<double*>(smem_raw + L.dkeys);                           // float64 key, by pedestrian index
    float4* pobs = reinterpret_cast<float4*>(smem_raw + L.pobs);                             // px, py, vx, vy in the robot frame (float32 like PedInfo)
    float* row = reinterpret_cast<float*>(smem_raw + L.row);                                 // this robot's ped_vector_states row
    __shared__ Tf2 s_world_base;
    __shared__ int s_npaint;
    if (tid == 0) s_npaint = 0;
    if (tid == 0) s_world_base = tf_inv(d.vconst[idx].base_world);      // (k_view_consts ran before the fork)
    const int npm = c.img * c.img;
    if ((npm & 3) == 0) for (int k = tid; k < npm / 4; k += PED_THREADS) reinterpret_cast<int4*>(winner)[k] = make_int4(-1, -1, -1, -1);
    else for (int k = tid; k < npm; k += PED_THREADS) winner[k] = -1;
    __syncthreads();
    const Tf2 world_base = s_world_base;
    for (int k = tid; k < (c.pvs_len + 3) / 4; k += PED_THREADS)      // (the row's slot is padded to 16 bytes)
        reinterpret_cast<float4*>(row)[k] = make_float4(k == 0 ? (float)c.P : 0.f, 0.f, 0.f, 0.f);
    for (int j = tid; j < L.n_sort; j += PED_THREADS) {
        unsigned fk = 0x7F800000u;        // +inf: padding sorts last
        if (j < c.P) {
            const int pi = s * c.P + j;
            double bx, by, bvx, bvy;
            tf_apply(world_base, PDF(d, PD_X, pi), PDF(d, PD_Y, pi), bx, by);
            tf_rotate(world_base, PDF(d, PD_VX, pi), PDF(d, PD_VY, pi), bvx, bvy);
            const float px = (float)bx, py = (float)by;
            pobs[j] = make_float4(px, py, (float)bvx, (float)bvy);
            const double dk = (double)px * (double)px + (double)py * (double)py;   // python: float(x)**2 + float(y)**2
            dkeys[j] = dk;
            fk = __float_as_uint(__double2float_rn(dk));
        }
        keys[j] = ((unsigned long long)fk << 32) | (unsigned)j;
    }
    __syncthreads();
    // bitonic network on the packed words (all distinct: the index is part of the word)
    if (L.n_sort == 2 * PED_THREADS) {
        // 2 elements per thread in registers (2t, 2t+1): partners inside a warp are reached with shuffles, only the last
        // stages (partner thread >= 32 lanes away) go through shared memory
        unsigned long long k0 = keys[2 * tid], k1 = keys[2 * tid + 1];
        for (int k = 2; k <= 2 * PED_THREADS; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                const bool up = ((2 * tid) & k) == 0;
                if (jj == 1) {
                    if ((k0 > k1) == up) { const unsigned long long tk = k0; k0 = k1; k1 = tk; }
                    continue;
                }
                const int tj = jj >> 1;                       // partner thread = tid ^ tj holds the partners of both elements
                unsigned long long p0, p1;
                if (tj < 32) {
                    p0 = __shfl_xor_sync(0xffffffffu, k0, tj); p1 = __shfl_xor_sync(0xffffffffu, k1, tj);
                } else {
                    __syncthreads();
                    keys[2 * tid] = k0; keys[2 * tid + 1] = k1;
                    __syncthreads();
                    const int pt = tid ^ tj;
                    p0 = keys[2 * pt]; p1 = keys[2 * pt + 1];
                }
                const bool take_min = (((2 * tid) & jj) == 0) == up;
                if ((k0 > p0) == take_min) k0 = p0;
                if ((k1 > p1) == take_min) k1 = p1;
            }
        __syncthreads();
        keys[2 * tid] = k0; keys[2 * tid + 1] = k1;
    } else
    for (int k = 2; k <= L.n_sort; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int t = tid; t < (L.n_sort >> 1); t += PED_THREADS) {
                const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1)), p = i | jj;      // the pair (i, i + jj) of this stage
                const unsigned long long ka = keys[i], kb = keys[p];
                const bool up = (i & k) == 0;
                if ((ka > kb) == up) { keys[i] = kb; keys[p] = ka; }
            }
            __syncthreads();
        }
    {   // runs of equal float32 keys: order by (float64 key, index) like the stable python sort
        bool tie = false;
        __syncthreads();
        for (int i = tid; i + 1 < c.P; i += PED_THREADS) tie |= (keys[i] >> 32) == (keys[i + 1] >> 32);
        if (__syncthreads_or(tie)) {
            if (tid == 0)
                for (int i = 1; i < c.P; i++) {          // insertion inside each run
                    const unsigned long long w = keys[i];
                    const int ji = (int)(unsigned)w; const double di = dkeys[ji];
                    int p = i;
                    while (p > 0 && (keys[p - 1] >> 32) == (w >> 32)) {
                        const int jp = (int)(unsigned)keys[p - 1]; const double dp = dkeys[jp];
                        if (dp > di || (dp == di && jp > ji)) { keys[p] = keys[p - 1]; p--; } else break;
                    }
                    keys[p] = w;
                }
            __syncthreads();
        }
    }
    for (int q = tid; q < c.P; q += PED_THREADS) {       // q = rank (0 = nearest), j = pedestrian
        const int j = (int)(unsigned)keys[q];
        const float4 o = pobs[j];
        const double px = o.x, py = o.y;
        const double ped_r = d.ped_r_round[j];
        const float f5 = (float)ped_r, f6 = (float)(ped_r + ty.size_last), f7 = (float)sqrt(px * px + py * py);
        if (q < c.max_ped) {
            float* w = row + 1 + (size_t)q * c.ped_vec_dim;
            w[0] = o.x; w[1] = o.y; w[2] = o.z; w[3] = o.w; w[4] = f5; w[5] = f6; w[6] = f7;
        }
        if (q == 0) RBF(d, RB_MIND, idx) = (double)(f7 - f6);      // NearbyPed.set(i, ped_tmp[7] - ped_tmp[6]) in float32 (yaml_env.py:455-456)
        if (px > 3 || px < -3 || py > 3 || py < -3) continue;
        // inside the 6 m x 6 m ped map: its cell window (python floor divisions, yaml_env.py:414-415) goes to a list; the few
        // pedestrians that are painted are then shared out over the CTA instead of holding everybody up at the barrier
        const double tmx = -px + 3, tmy = -py + 3;
        const int x0 = (int)py_floordiv(tmx - c.ped_image_r, c.ped_res), x1 = (int)py_floordiv(tmx + c.ped_image_r, c.ped_res);
        const int y0 = (int)py_floordiv(tmy - c.ped_image_r, c.ped_res), y1 = (int)py_floordiv(tmy + c.ped_image_r, c.ped_res);
        const int e = atomicAdd(&s_npaint, 1);
        paint[e] = make_int4((q << 16) | j, (max(x0, 0) << 16) | max(min(x1, c.img), 0), (max(y0, 0) << 16) | max(min(y1, c.img), 0), 0);
    }
    __syncthreads();
    if (tid == 0) d.o_mind[idx] = (float)RBF(d, RB_MIND, idx);
    for (int e = tid >> 5; e < s_npaint; e += PED_THREADS / 32) {      // one warp per painted pedestrian, one lane per cell of its window
        const int4 pe = paint[e];
        const int j = pe.x & 0xFFFF, x0 = pe.y >> 16, x1 = pe.y & 0xFFFF, y0 = pe.z >> 16, y1 = pe.z & 0xFFFF;
        const int wdt = y1 - y0, ncell = (x1 - x0) * wdt;
        if (ncell <= 0) continue;
        const float4 o = pobs[j];
        const double tmx = -(double)o.x + 3, tmy = -(double)o.y + 3;
        for (int u = tid & 31; u < ncell; u += 32) {
            const int jj = x0 + u / wdt, kk = y0 + u % wdt;
            const double ddx = (jj + 0.5) * c.ped_res - tmx, ddy = (kk + 0.5) * c.ped_res - tmy;
            if (ddx * ddx + ddy * ddy < c.ped_image_r * c.ped_image_r) atomicMax(&winner[jj * c.img + kk], pe.x);   // farther pedestrians overwrite
        }
    }
    float* pvs = d.o_pvs + (size_t)idx * c.pvs_len;
    for (int k = tid; k < c.pvs_len; k += PED_THREADS) pvs[k] = row[k];
    __syncthreads();
    // 3 x img x img float32, mostly zeros: 16-byte stores (each robot's block starts on a 16-byte boundary when img*img*3 % 4 == 0)
    float* pm = d.o_pmap + (size_t)idx * 3 * npm;
    auto value = [&](int ch, int cell) -> float {
        const int wv = winner[cell];
        if (wv < 0) return 0.f;
        const float4 o = pobs[wv & 0xFFFF];
        return ch == 0 ? 1.0f : (ch == 1 ? o.z : o.w);
    };
    if ((npm & 3) == 0) {
        for (int v = tid; v < npm / 4; v += PED_THREADS) {      // four cells of all three channels per pass
            const int4 w4 = reinterpret_cast<const int4*>(winner)[v];
            float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, o2 = o0;
            if ((w4.x & w4.y & w4.z & w4.w) >= 0) {      // some cell is painted
                o0 = make_float4(value(0, 4 * v), value(0, 4 * v + 1), value(0, 4 * v + 2), value(0, 4 * v + 3));
                o1 = make_float4(value(1, 4 * v), value(1, 4 * v + 1), value(1, 4 * v + 2), value(1, 4 * v + 3));
                o2 = make_float4(value(2, 4 * v), value(2, 4 * v + 1), value(2, 4 * v + 2), value(2, 4 * v + 3));
            }
            reinterpret_cast<float4*>(pm)[v] = o0;
            reinterpret_cast<float4*>(pm + (size_t)npm)[v] = o1;
            reinterpret_cast<float4*>(pm + 2 * (size_t)npm)[v] = o2;
        }
    } else {
        for (int v = tid; v < 3 * npm; v += PED_THREADS) { const int ch = v / npm; pm[v] = value(ch, v - ch * npm); }
    }
}
