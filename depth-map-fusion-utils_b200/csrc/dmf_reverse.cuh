// dmf_reverse.cuh -- K2: reverse per-voxel visibility march (reverseRayTraceFast / reverseRayTrace,
// include/RayTracingEngine.hpp:45-226) and K3: the z-buffer splat (rayTraceVolume, :498-564), plus the small
// volume-preparation kernels.
#pragma once
#include "dmf_device.cuh"

// getHash for arbitrary (possibly out-of-range) coordinates: ints sign-extend into the xor (Volume.hpp:135-141)
__device__ __forceinline__ u64 hash_coords(int x, int y, int z) {
    u64 h = (u64)(long long)x;
    return (h << 40) ^ (u64)(long long)(int)((unsigned)y << 20) ^ (u64)(long long)z;
}

__device__ __forceinline__ u64 hash_point(const VolDev& v, float x, float y, float z, unsigned& n_exact) {
    int ix = voxel_index(x, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
    int iy = voxel_index(y, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
    int iz = voxel_index(z, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
    return hash_coords(ix, iy, iz);
}

// centroid of occupied voxel `occ` as reverseRayTraceFast builds it (:148-157):
//   float x = xid*xdelta_ + xmin_;  centroid = float(x + xdelta_/2.0)
__device__ __forceinline__ void occ_centroid(const VolDev& v, int occ, int& xid, int& yid, int& zid, float& cx, float& cy, float& cz) {
    const u64 id = __ldg(v.occ_ids + occ);
    xid = (int)(id >> 40); yid = (int)((id >> 20) & 0xFFFFFu); zid = (int)(id & 0xFFFFFu);
    float x = __double2float_rn(__dadd_rn(__dmul_rn((double)xid, v.delta[0]), v.vmin[0]));
    float y = __double2float_rn(__dadd_rn(__dmul_rn((double)yid, v.delta[1]), v.vmin[1]));
    float z = __double2float_rn(__dadd_rn(__dmul_rn((double)zid, v.delta[2]), v.vmin[2]));
    cx = __double2float_rn(__dadd_rn((double)x, v.half[0]));
    cy = __double2float_rn(__dadd_rn((double)y, v.half[1]));
    cz = __double2float_rn(__dadd_rn((double)z, v.half[2]));
}

__global__ void k_centroid_hash(const VolDev v, u64* out) {
    for (int occ = blockIdx.x * blockDim.x + threadIdx.x; occ < v.n_occ; occ += gridDim.x * blockDim.x) {
        int xid, yid, zid; float cx, cy, cz; unsigned ne = 0;
        occ_centroid(v, occ, xid, yid, zid, cx, cy, cz);
        out[occ] = hash_point(v, cx, cy, cz, ne);
    }
}

// exhaustive check of div1000 / div1000_short against __fdiv_rn over every float bit pattern
__global__ void k_selftest_div1000(unsigned long long* mismatches) {
    unsigned long long bad_long = 0, bad_short = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((unsigned)i);
        const float ref = __fdiv_rn(a, 1000.0f);
        const float l = div1000(a), sh = div1000_short(a);
        const bool nan_ref = ref != ref;
        if (nan_ref ? !(l != l) : (__float_as_uint(l) != __float_as_uint(ref))) {
            bad_long++;
            const unsigned mag = (unsigned)i & 0x7fffffffu;          // |a| as bits: track the range of failing magnitudes
            if (mag < 0x7f800000u) { atomicMax((unsigned*)(mismatches + 2), mag); atomicMin((unsigned*)(mismatches + 3), mag); }
        }
        if (nan_ref ? !(sh != sh) : (__float_as_uint(sh) != __float_as_uint(ref))) {
            bad_short++;
            const unsigned mag = (unsigned)i & 0x7fffffffu;
            if (mag < 0x7f800000u && mag > 0x0d000000u) atomicAdd(mismatches + 4, 1ull);   // failures of the short form above 2^-101
        }
    }
    if (bad_long) atomicAdd(mismatches, bad_long);
    if (bad_short) atomicAdd(mismatches + 1, bad_short);
}

// Eigen::Affine3f::inverse() (rule E5 of oracle/dmf_oracle.hpp): cofactor inverse * (1/det), translation = -(Linv*t)
__global__ void k_invert_poses(const float* __restrict__ poses, float* __restrict__ inv, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m[3][4];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) m[r][c] = poses[12 * (size_t)i + 4 * r + c];
    auto cof = [&](int a, int b) {
        int a1 = (a + 1) % 3, a2 = (a + 2) % 3, b1 = (b + 1) % 3, b2 = (b + 2) % 3;
        return __fsub_rn(__fmul_rn(m[a1][b1], m[a2][b2]), __fmul_rn(m[a1][b2], m[a2][b1]));
    };
    float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    float det = sum3(__fmul_rn(c0, m[0][0]), __fmul_rn(c1, m[1][0]), __fmul_rn(c2, m[2][0]));
    float invdet = __fdiv_rn(1.0f, det);
    float r[3][4];
    r[0][0] = __fmul_rn(c0, invdet); r[0][1] = __fmul_rn(c1, invdet); r[0][2] = __fmul_rn(c2, invdet);
    r[1][0] = __fmul_rn(cof(0, 1), invdet); r[1][1] = __fmul_rn(cof(1, 1), invdet); r[1][2] = __fmul_rn(cof(2, 1), invdet);
    r[2][0] = __fmul_rn(cof(0, 2), invdet); r[2][1] = __fmul_rn(cof(1, 2), invdet); r[2][2] = __fmul_rn(cof(2, 2), invdet);
    for (int a = 0; a < 3; a++)
        r[a][3] = sum3(__fmul_rn(-r[a][0], m[0][3]), __fmul_rn(-r[a][1], m[1][3]), __fmul_rn(-r[a][2], m[2][3]));
    for (int a = 0; a < 3; a++) for (int b = 0; b < 4; b++) inv[12 * (size_t)i + 4 * a + b] = r[a][b];
}

struct RevArgs {
    VolDev vol;
    AngleTest angle;
    const float* __restrict__ poses;       // [n_views][12]
    const float* __restrict__ inv_poses;   // [n_views][12]
    double fx, cx, fy, cy;
    int H, W;
    const u64* __restrict__ centroid_hash; // [n_occ]  (fast)
    const float* __restrict__ ax[3];       // float-accumulated axes (whole-grid variants)
    int nax[3];
    unsigned* vis;                         // emitted; row of view v at vis + v * vis_stride32
    unsigned* unocc;                       // not occluded; dense [n_views][vis_words32]
    int vis_words32;
    unsigned vis_stride32;                 // == vis_words32 unless the rows live in the gathered buffer of a sharded sweep
    PubTable pub;                          // sharded sweep (FAST only): push finished rows to the peer GPUs
    int* found_any;
    int viz;
    int* view_mark;
    unsigned* good_bits;
    u64* emit_list;                        // whole-grid variant: [n_views][emit_cap][2] = (scan index, centroid hash)
    unsigned* emit_count;                  // [n_views]
    unsigned emit_cap;
    int* zbuf;                             // [H*W] (z-buffer)
    u64* counters;
    int step_cap;
    unsigned pnyz, bias;                   // pdim_y * pdim_z;  0x4B400000 * (pnyz + pdim_z + 1): the three shifter offsets of a line probe's index, folded on the host
};

__device__ __forceinline__ void camera_pixel(const RevArgs& a, float xx, float yy, float zz, int& r, int& c) {   // deProjectPoint, Camera.hpp:32-38
    c = to_int_x86(round(__dadd_rn(__ddiv_rn(__dmul_rn((double)xx, a.fx), (double)zz), a.cx)));
    r = to_int_x86(round(__dadd_rn(__ddiv_rn(__dmul_rn((double)yy, a.fy), (double)zz), a.cy)));
}

// The 1 mm march from `centroid` towards (and past) the camera (:81-103, :172-200).  true = occluded.
// FMT 0: bit grid, every step evaluated.  FMT 1: distance bytes -- a step that lands in a voxel with distance byte
// d >= 2 proves that the next floor((d-1.25)/step) steps land in empty voxels (they cannot collide, cannot be the origin voxel;
// they cannot leave the volume because the line loop cuts every jump at the end of its slab interval and the exact step folds
// the border distance in, byte_with_border), so they are counted and skipped; results and counters are unchanged.
// Samples are c + (v*depth)/1000 in float: within 3 roundings (< rev_eps voxels) of the line c + depth*(v/1000).
template <int FMT>
__device__ __forceinline__ bool march_collides(const RevArgs& a, float cx, float cy, float cz, float vx, float vy, float vz, u64 chash, int d0,
                                               unsigned& n_samples, unsigned& n_inb, unsigned& n_exact, unsigned& n_runaway,
                                               unsigned& n_f64, unsigned& n_skip) {
    const VolDev& v = a.vol;
    const float lo0 = v.lo[0], lo1 = v.lo[1], lo2 = v.lo[2], hi0 = v.hi[0], hi1 = v.hi[1], hi2 = v.hi[2];
    const float in0 = v.inv32[0], in1 = v.inv32[1], in2 = v.inv32[2], cc0 = v.c32[0], cc1 = v.c32[1], cc2 = v.c32[2];
    const float er0 = v.err32[0], er1 = v.err32[1], er2 = v.err32[2];
    const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
    // the fast a/1000 is exact only for |a| >= 2^-100 (dmf_device.cuh); a = v_i * depth with depth >= 1
    const bool fast_div = fminf(fabsf(vx), fminf(fabsf(vy), fabsf(vz))) >= 7.888609052210118e-31f;
    const float step = fmaxf(fabsf(vx) * fabsf(in0), fmaxf(fabsf(vy) * fabsf(in1), fabsf(vz) * fabsf(in2))) * 0.001f;   // voxels per 1 mm step
    const float rq = 0.999999f / fmaxf(step, 1e-3f);                   // rounded down: a jump computed from it never over-skips
    const bool skip_ok = FMT == 1 && fmaxf(v.rev_eps[0], fmaxf(v.rev_eps[1], v.rev_eps[2])) <= 0.1f;
    const float kM = 12582912.0f;
    int depth = d0;
    float s = (float)d0;
    // ---- line-first phase (FMT 1): follow Q(s) = QC + s*QV in voxel units and let the distance byte of the line point's voxel
    // decide, as k_forward_line does.  d >= 2: this step and the next floor((d-1.25)/step) are in-bounds, not the origin voxel,
    // not occupied.  d <= 1 with the line point >= rev_esafe away from every voxel face: the reference's sample is in the SAME
    // voxel, so d = 1 is a plain in-bounds step and d = 0 is either the origin voxel (continue) or an occluder (return true).
    // Anything else -- near a face, or outside the slab where the line is >= 0.25 voxel inside [0, min(ext, dim)] -- falls
    // through to the exact step below.  No output of the reverse march needs an exact position, so most marches never do.
    float qc0 = 0.f, qc1 = 0.f, qc2 = 0.f, qv0 = 0.f, qv1 = 0.f, qv2 = 0.f;
    int s_in = 1, s_out = 0;
    if (skip_ok) {
        qc0 = fmaf(cx, in0, cc0); qc1 = fmaf(cy, in1, cc1); qc2 = fmaf(cz, in2, cc2);
        qv0 = (vx * 0.001f) * in0; qv1 = (vy * 0.001f) * in1; qv2 = (vz * 0.001f) * in2;
        float t0 = -1e30f, t1 = 1e30f;
        const float qc[3] = {qc0, qc1, qc2}, qv[3] = {qv0, qv1, qv2};
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const float r = fabsf(qv[ax]) > 1e-12f ? __fdividef(1.0f, qv[ax]) : 1e30f;
            const float ta = (0.25f - qc[ax]) * r, tb = (fminf(v.ext[ax], (float)v.dim[ax]) - 0.25f - qc[ax]) * r;
            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
        }
        if (t0 <= t1) {
            // t0, t1 carry a relative error < 2^-20 (one subtraction, __fdividef, one product): guard g(t) = 2^-6 + |t| * 2^-17 steps on each
            // end, as in k_tile_start.  (A whole step of guard per end, as in round 1, puts one more exactly evaluated step into the band
            // where an unoccluded ray leaves the volume -- and those bands are most of the exact steps of a sweep.)
            s_in = (int)fminf(fmaxf(ceilf(t0 + fmaf(fabsf(t0), 7.62939453125e-06f, 0.015625f)), (float)d0), 1.0e9f);
            s_out = (int)fminf(fmaxf(floorf(t1 - fmaf(fabsf(t1), 7.62939453125e-06f, 0.015625f)), -1.0f), 1.0e9f);
        }
    }
    const float e_safe = v.rev_esafe;
    for (;;) {
        if (depth - d0 > a.step_cap) { n_runaway++; return false; }
        if (FMT == 1 && depth >= s_in && depth <= s_out) {
            // the line is followed up to s_out and never beyond the step cap (there the exact step below reports the runaway)
            const float s_outf = (float)s_out, s_limf = fminf(s_outf, (float)(d0 + a.step_cap));
            const float c1 = fmaf(-1.25f, rq, 1.0f);                                // this step + the skipped ones: floor(d * rq + c1)
            bool need_exact = false;
            for (;;) {
                const float q0 = fmaf(s, qv0, qc0), q1 = fmaf(s, qv1, qc1), q2 = fmaf(s, qv2, qc2);
                const float m0 = __fadd_rd(q0, kM), m1 = __fadd_rd(q1, kM), m2 = __fadd_rd(q2, kM);
                unsigned lidx = (unsigned)__float_as_int(m0) * a.pnyz + ((unsigned)__float_as_int(m1) * pnz + ((unsigned)__float_as_int(m2) - a.bias));
                DMF_CHECK_IDX(lidx, v.n_cells, a.counters);
                const unsigned d = __ldg(v.bytes + lidx);
                float adv = 1.0f;
                if (d >= 2u) adv = __fadd_rd(fmaf(__int_as_float(0x4B000000 | (int)d) - 8388608.0f, rq, c1), kM) - kM;
                else {
                    const float f0 = q0 - (m0 - kM), f1 = q1 - (m1 - kM), f2 = q2 - (m2 - kM);
                    if (!(fminf(f0, fminf(f1, f2)) >= e_safe && fmaxf(f0, fmaxf(f1, f2)) <= 1.0f - e_safe)) { need_exact = true; break; }
                    if (d == 0u && hash_coords(__float_as_int(m0) - 0x4B400000, __float_as_int(m1) - 0x4B400000, __float_as_int(m2) - 0x4B400000) != chash) {   // an occupied voxel other than the origin: occluded
                        const unsigned n = (unsigned)((int)s - depth);
                        n_samples += n + 1u; n_inb += n + 1u; n_skip += n + 1u;
                        return true;
                    }
                }
                s += adv;
                if (!(s <= s_limf)) break;
            }
            s = fminf(s, s_outf + 1.0f);            // the bytes know nothing of the boundary: steps up to s_out are in bounds by the slab test, later ones are not proven
            const int nd = (int)s;
            const unsigned n = (unsigned)(nd - depth);
            n_samples += n; n_inb += n; n_skip += n;
            depth = nd;
            if (!need_exact) continue;
        }
        // centroid + v*double(depth)/1000.0 in float (rule E4)
        const float ax = __fmul_rn(vx, s), ay = __fmul_rn(vy, s), az = __fmul_rn(vz, s);
        float qx, qy, qz;
        if (fast_div) { qx = div1000_short(ax); qy = div1000_short(ay); qz = div1000_short(az); }
        else { qx = __fdiv_rn(ax, 1000.0f); qy = __fdiv_rn(ay, 1000.0f); qz = __fdiv_rn(az, 1000.0f); }
        const float px = __fadd_rn(cx, qx), py = __fadd_rn(cy, qy), pz = __fadd_rn(cz, qz);
        n_samples++;
        if (!(px > lo0 && px < hi0 && py > lo1 && py < hi1 && pz > lo2 && pz < hi2)) return false;   // validPoints == false: break
        n_inb++;
        bool unsafe = false;
        int ix = voxel_index_f32(px, in0, cc0, er0, unsafe);
        int iy = voxel_index_f32(py, in1, cc1, er1, unsafe);
        int iz = voxel_index_f32(pz, in2, cc2, er2, unsafe);
        if (unsafe) {
            n_f64++;
            ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
            iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
            iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
        }
        int n = 0;
        if (hash_coords(ix, iy, iz) != chash) {                    // hash == centroid_hash: same voxel as the origin, continue
            if (!coords_valid(v, ix, iy, iz)) return false;        // validCoords == false: break
            const unsigned idx = ((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz;
            if (FMT == 0) {
                if ((__ldg(v.bits + (idx >> 5)) >> (idx & 31)) & 1u) return true;
            } else {
                const unsigned d = byte_with_border(v, __ldg(v.bytes + idx), ix, iy, iz);
                if (d == 0u) return true;
                if (skip_ok && d >= 2u) {
                    const float df = __int_as_float(0x4B000000 | (int)d) - 8388608.0f;
                    n = min(__float_as_int(__fadd_rd((df - 1.25f) * rq, kM)) - 0x4B400000, a.step_cap);
                    n_samples += (unsigned)n; n_inb += (unsigned)n; n_skip += (unsigned)n;
                }
            }
        }
        depth += n + 1; s += (float)(n + 1);
    }
}

// Warp-reduce the per-thread tallies (REDUX) and add the non-zero ones to one of the 256 counter replicas with result-less
// atomics.  Must be called by all 32 lanes.  The rarely non-zero tallies are reduced only if any lane has one.
__device__ __forceinline__ void flush_counters(u64* counters, unsigned n_samples, unsigned n_inb, unsigned n_hits, unsigned n_exact,
                                               unsigned n_oob, unsigned n_ties, unsigned n_runaway, unsigned n_f64 = 0, unsigned n_skip = 0) {
    const unsigned full = 0xffffffffu;
    const unsigned c0 = __reduce_add_sync(full, n_samples), c1 = __reduce_add_sync(full, n_inb), c2 = __reduce_add_sync(full, n_hits);
    const unsigned c9 = __reduce_add_sync(full, n_skip);
    const unsigned rare = __reduce_or_sync(full, n_exact | n_oob | n_ties | n_runaway | n_f64);
    unsigned c3 = 0, c4 = 0, c5 = 0, c7 = 0, c8 = 0;
    if (rare) {
        c3 = __reduce_add_sync(full, n_exact); c4 = __reduce_add_sync(full, n_oob); c5 = __reduce_add_sync(full, n_ties);
        c7 = __reduce_add_sync(full, n_runaway); c8 = __reduce_add_sync(full, n_f64);
    }
    if ((threadIdx.x & 31) == 0) {
        u64* const g = counter_slot(counters);
        if (c0) atomicAdd(g + 0, (unsigned long long)c0);
        if (c1) atomicAdd(g + 1, (unsigned long long)c1);
        if (c2) atomicAdd(g + 2, (unsigned long long)c2);
        if (c9) atomicAdd(g + 9, (unsigned long long)c9);
        if (c3) atomicAdd(g + 3, (unsigned long long)c3);
        if (c4) atomicAdd(g + 4, (unsigned long long)c4);
        if (c5) atomicAdd(g + 5, (unsigned long long)c5);
        if (c7) atomicAdd(g + 7, (unsigned long long)c7);
        if (c8) atomicAdd(g + 8, (unsigned long long)c8);
    }
}

// FAST = true : one thread per occupied voxel (grid.x covers n_occ), reverseRayTraceFast :136-226
// FAST = false: one thread per visited position of the float-accumulated whole-grid scan, reverseRayTrace :45-134
constexpr int REV_MIN_BLOCKS = 16;   // the march is a chain of dependent byte loads: 64 resident warps per SM measured 13 % faster than 32
template <bool FAST, int FMT>
__global__ void __launch_bounds__(128, REV_MIN_BLOCKS) k_reverse(const RevArgs a) {
    const VolDev& v = a.vol;
    const int view = blockIdx.y;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned n_samples = 0, n_inb = 0, n_hits = 0, n_exact = 0, n_oob = 0, n_ties = 0, n_runaway = 0, n_f64 = 0, n_skip = 0;
    const float* T = a.poses + 12 * (size_t)view;
    const float* I = a.inv_poses + 12 * (size_t)view;
    bool live = true, f_unocc = false, f_emit = false;
    int occ = -1; float cx = 0, cy = 0, cz = 0; u64 chash = 0;
    if (FAST) {
        live = t < (size_t)v.n_occ;
        if (live) {
            int xid, yid, zid;
            occ = v.rev_perm ? (int)__ldg(v.rev_perm + t) : (int)t;               // Morton work order (dmf_volume.cuh)
            occ_centroid(v, occ, xid, yid, zid, cx, cy, cz);
            chash = __ldg(a.centroid_hash + occ);
        }
    } else {
        const size_t total = (size_t)a.nax[0] * a.nax[1] * a.nax[2];
        live = t < total;
        if (live) {
            const int k = (int)(t % a.nax[2]); const size_t u = t / a.nax[2];
            const int j = (int)(u % a.nax[1]); const int i = (int)(u / a.nax[1]);
            const float x = __ldg(a.ax[0] + i), y = __ldg(a.ax[1] + j), z = __ldg(a.ax[2] + k);
            const int ix = voxel_index(x, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
            const int iy = voxel_index(y, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
            const int iz = voxel_index(z, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
            if (!coords_valid(v, ix, iy, iz)) { n_oob++; live = false; }       // voxels_[xid][yid][zid] out of range (UB)
            else if (!occupied<0>(v, ix, iy, iz)) live = false;
            else {
                occ = occupied_ordinal(v, ix, iy, iz);
                cx = __double2float_rn(__dadd_rn((double)x, v.half[0]));
                cy = __double2float_rn(__dadd_rn((double)y, v.half[1]));
                cz = __double2float_rn(__dadd_rn((double)z, v.half[2]));
                chash = hash_point(v, cx, cy, cz, n_exact);
            }
        }
    }
    if (live) {
        const float xx = affine_row(__ldg(I + 0), __ldg(I + 1), __ldg(I + 2), __ldg(I + 3), cx, cy, cz);
        const float yy = affine_row(__ldg(I + 4), __ldg(I + 5), __ldg(I + 6), __ldg(I + 7), cx, cy, cz);
        const float zz = affine_row(__ldg(I + 8), __ldg(I + 9), __ldg(I + 10), __ldg(I + 11), cx, cy, cz);
        int r, c;
        camera_pixel(a, xx, yy, zz, r, c);
        if (r >= 0 && r < a.H && c >= 0 && c < a.W) {                               // validPixel
            float vx = __fsub_rn(__ldg(T + 3), cx), vy = __fsub_rn(__ldg(T + 7), cy), vz = __fsub_rn(__ldg(T + 11), cz);
            const float n2 = sum3(__fmul_rn(vx, vx), __fmul_rn(vy, vy), __fmul_rn(vz, vz));
            if (n2 > 0.0f) { const float s = __fsqrt_rn(n2); vx = __fdiv_rn(vx, s); vy = __fdiv_rn(vy, s); vz = __fdiv_rn(vz, s); }
            const bool collided = march_collides<FMT>(a, cx, cy, cz, vx, vy, vz, chash, FAST ? 50 : 1, n_samples, n_inb, n_exact, n_runaway, n_f64, n_skip);
            if (!collided) {
                n_hits++;
                f_unocc = true;
                if (!FAST) {
                    if (a.found_any) raise_flag(a.found_any + view);
                    if (a.unocc) atomicOr(a.unocc + (size_t)view * a.vis_words32 + (occ >> 5), 1u << (occ & 31));
                }
                if (a.viz) a.view_mark[occ] = 1;                                     // :108, :205
                bool emit = false;
                if ((double)zz >= 0.20 && (double)zz <= 1.0) {                       // k_ZMin, k_ZMax (:109, :206)
                    emit = FAST ? any_normal_faces(v, a.angle, occ, vx, vy, vz, n_ties) : true;
                }
                f_emit = emit;
                if (emit && !FAST) {
                    if (a.viz) atomicOr(a.good_bits + (occ >> 5), 1u << (occ & 31)); // :112, :215
                    if (a.vis) atomicOr(a.vis + (size_t)view * a.vis_stride32 + (occ >> 5), 1u << (occ & 31));
                    if (a.emit_count) {
                        unsigned slot = atomicAdd(a.emit_count + view, 1u);
                        if (a.emit_list && slot < a.emit_cap) {
                            a.emit_list[2 * ((size_t)view * a.emit_cap + slot)] = (u64)t;
                            a.emit_list[2 * ((size_t)view * a.emit_cap + slot) + 1] = chash;
                        }
                    }
                }
            }
        }
    }
    if (FAST && v.rev_perm) {
        // Morton work order: a warp's 32 voxels are neighbours in space, not in ordinal -- every unoccluded voxel sets its own bits
        if (f_unocc) {
            if (a.found_any) raise_flag(a.found_any + view);
            const unsigned w = (unsigned)occ >> 5, m = 1u << (occ & 31);
            if (a.unocc) atomicOr(a.unocc + (size_t)view * a.vis_words32 + w, m);
            if (f_emit) {
                if (a.viz) atomicOr(a.good_bits + w, m);                             // :112, :215
                if (a.vis) atomicOr(a.vis + (size_t)view * a.vis_stride32 + w, m);
            }
        }
    } else if (FAST) {
        // thread t <-> occupied ordinal t: the 32 lanes of a warp own exactly one word of every per-voxel bitset, so the warp
        // publishes its results with one atomic per bitset instead of one per voxel
        const unsigned mu = __ballot_sync(0xffffffffu, f_unocc), me = __ballot_sync(0xffffffffu, f_emit);
        if ((threadIdx.x & 31) == 0 && mu) {
            if (a.found_any) raise_flag(a.found_any + view);
            if (a.unocc) atomicOr(a.unocc + (size_t)view * a.vis_words32 + (size_t)(t >> 5), mu);
            if (me) {
                if (a.viz) atomicOr(a.good_bits + (t >> 5), me);                     // :112, :215
                if (a.vis) atomicOr(a.vis + (size_t)view * a.vis_stride32 + (size_t)(t >> 5), me);
            }
        }
    }
    flush_counters(a.counters, n_samples, n_inb, n_hits, n_exact, n_oob, n_ties, n_runaway, n_f64, n_skip);
    if (FAST && a.pub.enabled) publish_view_row(a.pub, reinterpret_cast<u64*>(a.vis + (size_t)view * a.vis_stride32), view, gridDim.x);
}

// ---- K2 with a ray pool: reverseRayTraceFast on distance bytes (DMF_REVERSE_POOL=1; measured slower than k_reverse, see the
// launch site in dmf_abi_rest.cuh -- kept as the recorded alternative of VERDICT r1 item 6) ------------------------------------
// k_reverse<true,1> gives every occupied voxel a thread for its whole life: prologue (centroid, inverse transform, double-precision
// deProjectPoint: 56 % of the voxels of a shell scene end here, outside the image), march, epilogue (normal test).  The warp then
// waits for its longest march: ncu (profiles/r01_reverse_ncu_summary.txt) measured 14.8 of 32 lanes active, 11 in the line loop
// that carries 61 % of the instructions.  Here a block takes RP_ROUNDS x RP_ROUND consecutive voxels and works in phases:
//   P1  (all lanes busy)  every thread runs the prologue of its voxels of the round and pushes the rays that survive it -- with
//       their line set-up and slab interval already computed -- into a queue in shared memory;
//   P2  (the march)       lanes PULL rays from the queue: a lane whose march ends fetches the next ray at once, so the warp keeps
//       stepping with (nearly) all lanes until the queue is dry; then the block refills it (P1 of the next round) while the lanes
//       keep their unfinished rays in registers.  Only the last round drains;
//   P3  (all lanes busy)  the voxels whose ray reached the volume boundary unoccluded get the depth window and the normal test.
// The per-voxel results are collected in shared-memory bitsets (the block owns its 64 words of every row) and leave as plain
// 32-bit stores.  The march itself is march_collides, one probe per loop iteration; samples / inbounds / skipped are counted
// probe for probe as there, so the counters (and of course the bitsets) are identical to k_reverse<true,*>.
#ifndef DMF_RP_MIN_BLOCKS
#define DMF_RP_MIN_BLOCKS 8
#endif
#ifndef DMF_RP_ROUNDS
#define DMF_RP_ROUNDS 4
#endif
constexpr int RP_THREADS = 128, RP_ROUND = 512, RP_ROUNDS = DMF_RP_ROUNDS, RP_BLOCK_VOX = RP_ROUND * RP_ROUNDS;
struct RevRay { float cx, cy, cz, vx, vy, vz; int s_in, s_out; unsigned occ, pad; };

__global__ void __launch_bounds__(RP_THREADS, DMF_RP_MIN_BLOCKS) k_reverse_pool(const RevArgs a) {
    __shared__ RevRay s_q[RP_ROUND];
    __shared__ int s_head, s_tail;
    __shared__ unsigned s_unocc[RP_BLOCK_VOX / 32], s_emit[RP_BLOCK_VOX / 32];
    const VolDev& v = a.vol;
    const int view = blockIdx.y;
    const unsigned vox0 = blockIdx.x * (unsigned)RP_BLOCK_VOX;
    const int lane = threadIdx.x & 31;
    const float* T = a.poses + 12 * (size_t)view;
    const float* I = a.inv_poses + 12 * (size_t)view;
    for (int w = threadIdx.x; w < RP_BLOCK_VOX / 32; w += RP_THREADS) { s_unocc[w] = 0u; s_emit[w] = 0u; }
    if (threadIdx.x == 0) { s_head = 0; s_tail = 0; }
    unsigned n_samples = 0, n_inb = 0, n_hits = 0, n_exact = 0, n_ties = 0, n_runaway = 0, n_f64 = 0, n_skip = 0;

    const float lo0 = v.lo[0], lo1 = v.lo[1], lo2 = v.lo[2], hi0 = v.hi[0], hi1 = v.hi[1], hi2 = v.hi[2];
    const float in0 = v.inv32[0], in1 = v.inv32[1], in2 = v.inv32[2], cc0 = v.c32[0], cc1 = v.c32[1], cc2 = v.c32[2];
    const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
    const bool skip_ok = fmaxf(v.rev_eps[0], fmaxf(v.rev_eps[1], v.rev_eps[2])) <= 0.1f;
    const float kM = 12582912.0f, e_safe = v.rev_esafe;
    const int d0 = 50;                                              // the march of reverseRayTraceFast starts at 50 mm (:172)
    const float tx = __ldg(T + 3), ty = __ldg(T + 7), tz = __ldg(T + 11);

    // the ray a lane is marching (kept across refills)
    bool have = false, on_line = false, need_exact = false;           // need_exact: the current sample lies next to a voxel face (line inconclusive)
    float cx = 0, cy = 0, cz = 0, vx = 0, vy = 0, vz = 0, qc0 = 0, qc1 = 0, qc2 = 0, qv0 = 0, qv1 = 0, qv2 = 0, sf = 0, rq = 0, s_inf = 1.0f, s_outf = 0.0f;
    unsigned occ = 0; u64 chash = 0; bool fast_div = true;

    for (int round = 0; round < RP_ROUNDS; round++) {
        __syncthreads();                                            // the queue is dry (or this is the first round): refill from the start
        if (threadIdx.x == 0) { s_head = 0; s_tail = 0; }
        __syncthreads();
        // ---- P1: prologue of this round's voxels, all lanes busy ----
        for (int j = 0; j < RP_ROUND / RP_THREADS; j++) {
            const unsigned o = vox0 + (unsigned)(round * RP_ROUND + j * RP_THREADS) + threadIdx.x;
            bool live = o < (unsigned)v.n_occ;
            RevRay r;
            if (live) {
                int xid, yid, zid;
                occ_centroid(v, (int)o, xid, yid, zid, r.cx, r.cy, r.cz);
                const float xx = affine_row(__ldg(I + 0), __ldg(I + 1), __ldg(I + 2), __ldg(I + 3), r.cx, r.cy, r.cz);
                const float yy = affine_row(__ldg(I + 4), __ldg(I + 5), __ldg(I + 6), __ldg(I + 7), r.cx, r.cy, r.cz);
                const float zz = affine_row(__ldg(I + 8), __ldg(I + 9), __ldg(I + 10), __ldg(I + 11), r.cx, r.cy, r.cz);
                int pr, pc;
                camera_pixel(a, xx, yy, zz, pr, pc);
                live = pr >= 0 && pr < a.H && pc >= 0 && pc < a.W;                          // validPixel
            }
            if (live) {
                r.vx = __fsub_rn(tx, r.cx); r.vy = __fsub_rn(ty, r.cy); r.vz = __fsub_rn(tz, r.cz);
                const float n2 = sum3(__fmul_rn(r.vx, r.vx), __fmul_rn(r.vy, r.vy), __fmul_rn(r.vz, r.vz));
                if (n2 > 0.0f) { const float sn = __fsqrt_rn(n2); r.vx = __fdiv_rn(r.vx, sn); r.vy = __fdiv_rn(r.vy, sn); r.vz = __fdiv_rn(r.vz, sn); }
                // the slab in which the line is >= 0.25 voxel inside [0, min(ext, dim)] (as march_collides)
                r.s_in = 1; r.s_out = 0;
                if (skip_ok) {
                    const float c0 = fmaf(r.cx, in0, cc0), c1 = fmaf(r.cy, in1, cc1), c2 = fmaf(r.cz, in2, cc2);
                    const float w0 = (r.vx * 0.001f) * in0, w1 = (r.vy * 0.001f) * in1, w2 = (r.vz * 0.001f) * in2;
                    float t0 = -1e30f, t1 = 1e30f;
                    const float qc[3] = {c0, c1, c2}, qv[3] = {w0, w1, w2};
#pragma unroll
                    for (int ax = 0; ax < 3; ax++) {
                        const float rr = fabsf(qv[ax]) > 1e-12f ? __fdividef(1.0f, qv[ax]) : 1e30f;
                        const float ta = (0.25f - qc[ax]) * rr, tb = (fminf(v.ext[ax], (float)v.dim[ax]) - 0.25f - qc[ax]) * rr;
                        t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                    }
                    if (t0 <= t1) {
                        r.s_in = (int)fminf(fmaxf(ceilf(t0) + 1.0f, (float)d0), 1.0e9f);
                        r.s_out = (int)fminf(fmaxf(floorf(t1) - 1.0f, -1.0f), 1.0e9f);
                    }
                }
                r.occ = o; r.pad = 0u;
            }
            // warp-aggregated push
            const unsigned m = __ballot_sync(0xffffffffu, live);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_tail, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) s_q[base + __popc(m & ((1u << lane) - 1u))] = r;
        }
        __syncthreads();
        const int tail = s_tail;
        const bool last_round = round == RP_ROUNDS - 1;
        bool dry = false;                                                           // this lane has seen the queue empty in this round
        // ---- P2: the march; lanes pull rays ----
        // Two nested loops.  The inner one is the line loop of march_collides and nothing else (one vote + ~25 instructions per
        // probe); it runs for as long as at least `quorum` lanes can follow their line.  Everything rare -- fetching the next
        // ray, an exact step (outside the slab, or next to a voxel face), ending a ray -- happens in the outer loop, where the
        // lanes that do not need it wait; the vote bounds how many can be waiting while the inner loop runs.
        for (;;) {
            if (have && !on_line) {
                // ---- one exact step of march_collides (or the end of the ray) ----
                int result = -1;                                                     // -1: goes on, 0: reached the boundary unoccluded, 1: occluded
                if (sf - (float)d0 > (float)a.step_cap) { n_runaway++; result = 0; }
                else {
                    const float ax = __fmul_rn(vx, sf), ay = __fmul_rn(vy, sf), az = __fmul_rn(vz, sf);
                    float qx, qy, qz;
                    if (fast_div) { qx = div1000_short(ax); qy = div1000_short(ay); qz = div1000_short(az); }
                    else { qx = __fdiv_rn(ax, 1000.0f); qy = __fdiv_rn(ay, 1000.0f); qz = __fdiv_rn(az, 1000.0f); }
                    const float px = __fadd_rn(cx, qx), py = __fadd_rn(cy, qy), pz = __fadd_rn(cz, qz);
                    n_samples++;
                    if (!(px > lo0 && px < hi0 && py > lo1 && py < hi1 && pz > lo2 && pz < hi2)) result = 0;      // validPoints == false: break
                    else {
                        n_inb++;
                        bool unsafe = false;
                        int ix = voxel_index_f32(px, in0, cc0, v.err32[0], unsafe);
                        int iy = voxel_index_f32(py, in1, cc1, v.err32[1], unsafe);
                        int iz = voxel_index_f32(pz, in2, cc2, v.err32[2], unsafe);
                        if (unsafe) {
                            n_f64++;
                            ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
                            iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
                            iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
                        }
                        int n = 0;
                        if (hash_coords(ix, iy, iz) != chash) {                    // hash == centroid_hash: same voxel as the origin, continue
                            if (!coords_valid(v, ix, iy, iz)) result = 0;          // validCoords == false: break
                            else {
                                const unsigned d = byte_with_border(v, __ldg(v.bytes + (((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz)), ix, iy, iz);
                                if (d == 0u) result = 1;
                                else if (skip_ok && d >= 2u) {
                                    const float df = __int_as_float(0x4B000000 | (int)d) - 8388608.0f;
                                    n = min(__float_as_int(__fadd_rd((df - 1.25f) * rq, kM)) - 0x4B400000, a.step_cap);
                                    n_samples += (unsigned)n; n_inb += (unsigned)n; n_skip += (unsigned)n;
                                }
                            }
                        }
                        if (result < 0) sf += (float)(n + 1);
                    }
                }
                need_exact = false;
                if (result >= 0) {
                    if (result == 0) { const unsigned lv = occ - vox0; atomicOr(&s_unocc[lv >> 5], 1u << (lv & 31)); n_hits++; }
                    have = false;
                }
            }
            if (!have && !dry) {
                const int i = atomicAdd(&s_head, 1);
                dry = i >= tail;
                if (!dry) {
                    const RevRay r = s_q[i];
                    cx = r.cx; cy = r.cy; cz = r.cz; vx = r.vx; vy = r.vy; vz = r.vz; occ = r.occ;
                    // the line is followed for s in [s_in, s_out] and never beyond the step cap (there the exact step reports the runaway)
                    s_inf = (float)r.s_in; s_outf = fminf((float)r.s_out, (float)(d0 + a.step_cap));
                    chash = __ldg(a.centroid_hash + occ);
                    qc0 = fmaf(cx, in0, cc0); qc1 = fmaf(cy, in1, cc1); qc2 = fmaf(cz, in2, cc2);
                    qv0 = (vx * 0.001f) * in0; qv1 = (vy * 0.001f) * in1; qv2 = (vz * 0.001f) * in2;
                    fast_div = fminf(fabsf(vx), fminf(fabsf(vy), fabsf(vz))) >= 7.888609052210118e-31f;
                    const float step = fmaxf(fabsf(vx) * fabsf(in0), fmaxf(fabsf(vy) * fabsf(in1), fabsf(vz) * fabsf(in2))) * 0.001f;
                    rq = 1.0f / fmaxf(step, 1e-3f);
                    sf = (float)d0;
                    have = true; need_exact = false;
                }
            }
            const unsigned busy = __ballot_sync(0xffffffffu, have);
            if (busy == 0u) break;                                                  // nothing in flight in this warp and the queue is dry
            if (!last_round && busy != 0xffffffffu) break;                          // the queue is dry: refill it, in-flight rays stay in registers
            on_line = have && !need_exact && skip_ok && sf >= s_inf && sf <= s_outf;
            const int quorum = max(1, (3 * __popc(busy)) >> 2);                      // keep at least 3/4 of the lanes that hold a ray stepping
            float adv_sum = 0.0f;                                                    // samples this lane advanced in this run of the line loop
            while (__popc(__ballot_sync(0xffffffffu, on_line)) >= quorum) {
                if (on_line) {
                    const float q0 = fmaf(sf, qv0, qc0), q1 = fmaf(sf, qv1, qc1), q2 = fmaf(sf, qv2, qc2);
                    const float m0 = __fadd_rd(q0, kM), m1 = __fadd_rd(q1, kM), m2 = __fadd_rd(q2, kM);
                    const int ix = __float_as_int(m0) - 0x4B400000, iy = __float_as_int(m1) - 0x4B400000, iz = __float_as_int(m2) - 0x4B400000;
                    unsigned lidx = ((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz;
                    DMF_CHECK_IDX(lidx, v.n_cells, a.counters);
                    const unsigned d = __ldg(v.bytes + lidx);
                    if (d >= 2u) {
                        const float df = __int_as_float(0x4B000000 | (int)d) - 8388608.0f;
                        const float adv = fminf(fminf(__fadd_rd(fmaf(df - 1.25f, rq, 1.0f), kM) - kM, (float)a.step_cap + 1.0f), s_outf + 1.0f - sf);   // (never past the slab interval)
                        sf += adv; adv_sum += adv;
                        on_line = sf <= s_outf;
                    } else {
                        const float f0 = q0 - (m0 - kM), f1 = q1 - (m1 - kM), f2 = q2 - (m2 - kM);
                        if (fminf(f0, fminf(f1, f2)) >= e_safe && fmaxf(f0, fmaxf(f1, f2)) <= 1.0f - e_safe) {
                            adv_sum += 1.0f;                                         // resolved on the line: an in-bounds step
                            if (d == 0u && hash_coords(ix, iy, iz) != chash) { have = false; on_line = false; }      // an occupied voxel other than the origin: occluded
                            else { sf += 1.0f; on_line = sf <= s_outf; }
                        } else { on_line = false; need_exact = true; }              // next to a face: this sample takes the exact step
                    }
                }
            }
            { const unsigned n = (unsigned)adv_sum; n_samples += n; n_inb += n; n_skip += n; }
            // whoever still holds a ray either stays on its line (it only lost the vote) or takes an exact step next
            on_line = have && !need_exact && skip_ok && sf >= s_inf && sf <= s_outf;
        }
    }
    __syncthreads();
    // ---- P3: depth window + normal test of the unoccluded voxels; results out ----
    for (int j = 0; j < RP_BLOCK_VOX / RP_THREADS; j++) {
        const unsigned lv = (unsigned)(j * RP_THREADS) + threadIdx.x, o = vox0 + lv;
        if (!((s_unocc[lv >> 5] >> (lv & 31)) & 1u)) continue;
        int xid, yid, zid; float ccx, ccy, ccz;
        occ_centroid(v, (int)o, xid, yid, zid, ccx, ccy, ccz);
        const float zz = affine_row(__ldg(I + 8), __ldg(I + 9), __ldg(I + 10), __ldg(I + 11), ccx, ccy, ccz);
        if (a.viz) a.view_mark[o] = 1;                                                   // :205
        if ((double)zz >= 0.20 && (double)zz <= 1.0) {                                   // k_ZMin, k_ZMax (:206)
            float dx = __fsub_rn(tx, ccx), dy = __fsub_rn(ty, ccy), dz = __fsub_rn(tz, ccz);
            const float n2 = sum3(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz));
            if (n2 > 0.0f) { const float sn = __fsqrt_rn(n2); dx = __fdiv_rn(dx, sn); dy = __fdiv_rn(dy, sn); dz = __fdiv_rn(dz, sn); }
            if (any_normal_faces(v, a.angle, (int)o, dx, dy, dz, n_ties)) atomicOr(&s_emit[lv >> 5], 1u << (lv & 31));
        }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < RP_BLOCK_VOX / 32; w += RP_THREADS) {
        const unsigned gw = blockIdx.x * (unsigned)(RP_BLOCK_VOX / 32) + (unsigned)w;
        if (gw * 32u >= (unsigned)v.n_occ) break;
        const unsigned mu = s_unocc[w], me = s_emit[w];
        if (mu && a.found_any) raise_flag(a.found_any + view);
        if (a.unocc && mu) a.unocc[(size_t)view * a.vis_words32 + gw] = mu;              // the block owns these words of the (zeroed) rows
        if (me) {
            if (a.viz) atomicOr(a.good_bits + gw, me);                                   // :215 (shared by all views)
            if (a.vis) a.vis[(size_t)view * a.vis_stride32 + gw] = me;
        }
    }
    flush_counters(a.counters, n_samples, n_inb, n_hits, n_exact, 0u, n_ties, n_runaway, n_f64, n_skip);
    if (a.pub.enabled) publish_view_row(a.pub, reinterpret_cast<u64*>(a.vis + (size_t)view * a.vis_stride32), view, gridDim.x);
}

// ---- K7: willCollide (tests/CameraPathGen.cpp:128-156 and the unguarded copies in CameraMotionTSP.cpp:236-261,
// CameraMotionPlanner.cpp:246-271), one thread per segment a->b: 1 mm float march `a + v*double(depth)/1000.0` while
// depth <= |a-b|*1000; samples outside the volume are skipped (continue), an occupied voxel ends the march.
// FMT 1 skips provably empty in-bounds steps through the distance bytes exactly like march_collides.
template <int FMT>
__global__ void __launch_bounds__(128) k_segments_collide(const VolDev v, const float* __restrict__ A, const float* __restrict__ B, int n,
                                                          int guard_coords, unsigned char* __restrict__ out, u64* counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_samples = 0, n_inb = 0, n_hits = 0, n_exact = 0, n_f64 = 0, n_skip = 0;
    if (i < n) {
        const float ax = A[3 * i], ay = A[3 * i + 1], az = A[3 * i + 2], bx = B[3 * i], by = B[3 * i + 1], bz = B[3 * i + 2];
        const float ex = __fsub_rn(ax, bx), ey = __fsub_rn(ay, by), ez = __fsub_rn(az, bz);
        const double distance = (double)__fsqrt_rn(sum3(__fmul_rn(ex, ex), __fmul_rn(ey, ey), __fmul_rn(ez, ez)));   // (a-b).norm()
        float vx = __fsub_rn(bx, ax), vy = __fsub_rn(by, ay), vz = __fsub_rn(bz, az);                                // (b-a).normalized()
        const float n2 = sum3(__fmul_rn(vx, vx), __fmul_rn(vy, vy), __fmul_rn(vz, vz));
        if (n2 > 0.0f) { const float s = __fsqrt_rn(n2); vx = __fdiv_rn(vx, s); vy = __fdiv_rn(vy, s); vz = __fdiv_rn(vz, s); }
        // `if(depth > distance*1000) break;`  <=>  run while depth <= floor(distance*1000)
        const double lim = floor(__dmul_rn(distance, 1000.0));
        const int dmax = lim >= 1.0e9 ? 1000000000 : (lim >= 0.0 ? (int)lim : 0);       // NaN/negative: no step passes the test
        const float lo0 = v.lo[0], lo1 = v.lo[1], lo2 = v.lo[2], hi0 = v.hi[0], hi1 = v.hi[1], hi2 = v.hi[2];
        const float in0 = v.inv32[0], in1 = v.inv32[1], in2 = v.inv32[2], cc0 = v.c32[0], cc1 = v.c32[1], cc2 = v.c32[2];
        const float er0 = v.err32[0], er1 = v.err32[1], er2 = v.err32[2];
        const unsigned pny = (unsigned)v.pdim[1], pnz = (unsigned)v.pdim[2];
        const bool fast_div = fminf(fabsf(vx), fminf(fabsf(vy), fabsf(vz))) >= 7.888609052210118e-31f;
        const float step = fmaxf(fabsf(vx) * fabsf(in0), fmaxf(fabsf(vy) * fabsf(in1), fabsf(vz) * fabsf(in2))) * 0.001f;
        const float rq = 1.0f / fmaxf(step, 1e-3f);
        const bool skip_ok = FMT == 1 && fmaxf(v.rev_eps[0], fmaxf(v.rev_eps[1], v.rev_eps[2])) <= 0.1f;
        const float kM = 12582912.0f;
        bool collided = false;
        int depth = 1;
        float s = 1.0f;
        while (depth <= dmax) {
            const float qx0 = __fmul_rn(vx, s), qy0 = __fmul_rn(vy, s), qz0 = __fmul_rn(vz, s);
            float qx, qy, qz;
            if (fast_div) { qx = div1000_short(qx0); qy = div1000_short(qy0); qz = div1000_short(qz0); }
            else { qx = __fdiv_rn(qx0, 1000.0f); qy = __fdiv_rn(qy0, 1000.0f); qz = __fdiv_rn(qz0, 1000.0f); }
            const float px = __fadd_rn(ax, qx), py = __fadd_rn(ay, qy), pz = __fadd_rn(az, qz);
            n_samples++;
            int nskip = 0;
            if (px > lo0 && px < hi0 && py > lo1 && py < hi1 && pz > lo2 && pz < hi2) {
                n_inb++;
                bool unsafe = false;
                int ix = voxel_index_f32(px, in0, cc0, er0, unsafe);
                int iy = voxel_index_f32(py, in1, cc1, er1, unsafe);
                int iz = voxel_index_f32(pz, in2, cc2, er2, unsafe);
                if (unsafe) {
                    n_f64++;
                    ix = voxel_index(px, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
                    iy = voxel_index(py, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
                    iz = voxel_index(pz, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
                }
                // index == dim (the unguarded copies read out of range there): the padded plane is empty either way
                if (!guard_coords || coords_valid(v, ix, iy, iz)) {
                    const unsigned idx = ((unsigned)ix * pny + (unsigned)iy) * pnz + (unsigned)iz;
                    if (FMT == 0) {
                        if ((__ldg(v.bits + (idx >> 5)) >> (idx & 31)) & 1u) { collided = true; break; }
                    } else {
                        const unsigned d = byte_with_border(v, __ldg(v.bytes + idx), ix, iy, iz);
                        if (d == 0u) { collided = true; break; }
                        if (skip_ok && d >= 2u) {
                            const float df = __int_as_float(0x4B000000 | (int)d) - 8388608.0f;
                            nskip = min(__float_as_int(__fadd_rd((df - 1.25f) * rq, kM)) - 0x4B400000, dmax - depth);
                            n_samples += (unsigned)nskip; n_inb += (unsigned)nskip; n_skip += (unsigned)nskip;
                        }
                    }
                }
            }
            depth += nskip + 1; s += (float)(nskip + 1);
        }
        out[i] = collided ? 1 : 0;
        n_hits = collided ? 1u : 0u;
    }
    flush_counters(counters, n_samples, n_inb, n_hits, n_exact, 0u, 0u, 0u, n_f64, n_skip);
}

// visibility bitset -> occupied ordinals in ascending order (= emission order of reverseRayTraceFast).
// Multi-block (a single view's 140 k voxels in one block took 15-29 us): blocks of BITS_TILE words count, one block
// per view scans the counts (k_win_offsets), the blocks emit.
constexpr int BITS_TILE = 256;                    // 32-bit words per block = threads per block (8192 voxels)
__global__ void __launch_bounds__(BITS_TILE) k_bits_count(const unsigned* __restrict__ vis, int vis_words32, int n_occ, unsigned* __restrict__ blk_cnt, int nb) {
    __shared__ unsigned s_warp[BITS_TILE / 32];
    const int view = blockIdx.y, w = blockIdx.x * BITS_TILE + threadIdx.x, nwords = (n_occ + 31) / 32;
    unsigned c = w < nwords ? (unsigned)__popc(vis[(size_t)view * vis_words32 + w]) : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int i = 0; i < BITS_TILE / 32; i++) t += s_warp[i]; blk_cnt[(size_t)view * nb + blockIdx.x] = t; }
}
__global__ void __launch_bounds__(BITS_TILE) k_bits_emit(const unsigned* __restrict__ vis, int vis_words32, int n_occ, const unsigned* __restrict__ blk_off, int nb,
                                                        int* __restrict__ out_occ, int stride) {
    __shared__ unsigned s_warp[BITS_TILE / 32];
    const int view = blockIdx.y, w = blockIdx.x * BITS_TILE + threadIdx.x, nwords = (n_occ + 31) / 32;
    unsigned m = w < nwords ? vis[(size_t)view * vis_words32 + w] : 0u;
    const unsigned c = (unsigned)__popc(m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = c;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    unsigned before = 0;
    for (int q = 0; q < warp; q++) before += s_warp[q];
    unsigned pos = blk_off[(size_t)view * nb + blockIdx.x] + before + x - c;
    while (m) { const int bit = __ffs(m) - 1; m &= m - 1; out_occ[(size_t)view * stride + pos++] = w * 32 + bit; }
}

// ---- K3: rayTraceVolume ---------------------------------------------------------------------------------------
// PASS 0: depth[r][c] = min(depth[r][c], int(round(zz*1000)))  (:528-534; -1 == "unset" is modelled as INT_MAX)
// PASS 1: voxel->view = 1 where depth[r][c] == d              (:559-562)
template <int PASS>
__global__ void __launch_bounds__(256) k_zbuffer(const RevArgs a, unsigned long long* n_splat, unsigned long long* n_minus_one) {
    const VolDev& v = a.vol;
    const size_t total = (size_t)a.nax[0] * a.nax[1] * a.nax[2];
    unsigned n_exact = 0;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % a.nax[2]); const size_t u = t / a.nax[2];
        const int j = (int)(u % a.nax[1]); const int i = (int)(u / a.nax[1]);
        const float x = __ldg(a.ax[0] + i), y = __ldg(a.ax[1] + j), z = __ldg(a.ax[2] + k);
        const int ix = voxel_index(x, v.vmin[0], v.delta[0], v.inv[0], v.c0[0], v.eps[0], n_exact);
        const int iy = voxel_index(y, v.vmin[1], v.delta[1], v.inv[1], v.c0[1], v.eps[1], n_exact);
        const int iz = voxel_index(z, v.vmin[2], v.delta[2], v.inv[2], v.c0[2], v.eps[2], n_exact);
        if (!coords_valid(v, ix, iy, iz)) continue;
        if (!occupied<0>(v, ix, iy, iz)) continue;
        const float cx = __double2float_rn(__dadd_rn((double)x, v.half[0]));
        const float cy = __double2float_rn(__dadd_rn((double)y, v.half[1]));
        const float cz = __double2float_rn(__dadd_rn((double)z, v.half[2]));
        const float* I = a.inv_poses;
        const float xx = affine_row(__ldg(I + 0), __ldg(I + 1), __ldg(I + 2), __ldg(I + 3), cx, cy, cz);
        const float yy = affine_row(__ldg(I + 4), __ldg(I + 5), __ldg(I + 6), __ldg(I + 7), cx, cy, cz);
        const float zz = affine_row(__ldg(I + 8), __ldg(I + 9), __ldg(I + 10), __ldg(I + 11), cx, cy, cz);
        int r, c;
        camera_pixel(a, xx, yy, zz, r, c);
        if (!(r >= 0 && r < a.H && c >= 0 && c < a.W)) continue;
        const float dm = roundf(__fmul_rn(zz, 1000.0f));
        const int d = (dm > -2147483904.0f && dm < 2147483648.0f) ? (int)dm : (int)0x80000000;
        if (PASS == 0) {
            atomicMin(a.zbuf + (size_t)r * a.W + c, d);
            atomicAdd(n_splat, 1ull);
            if (d == -1) atomicAdd(n_minus_one, 1ull);
        } else if (a.zbuf[(size_t)r * a.W + c] == d) {
            a.view_mark[occupied_ordinal(v, ix, iy, iz)] = 1;
        }
    }
}

__global__ void k_finish_zbuf(int* zbuf, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (zbuf[i] == 0x7fffffff) zbuf[i] = -1;
}
