// dmf_device.cuh -- device-side data layout and exact-arithmetic helpers shared by all kernels.
//
// Everything here must reproduce the reference's IEEE-754 results bit for bit, so the file is compiled
// with -fmad=false (no implicit contraction), default -prec-div/-prec-sqrt/-ftz=false, and every op whose
// rounding matters is spelled with an explicit round-to-nearest intrinsic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;

// The DMF_CNT_* counters are kept in DMF_COUNTER_SLOTS replicas, one 128-byte line each, and summed by dmf_counters():
// millions of blocks adding to four shared addresses would serialise in one L2 slice (same-address atomics retire at
// about one per clock), which at ~5 M atomics per launch was a measurable part of the march kernels' duration.
#define DMF_COUNTER_SLOTS 256
#define DMF_COUNTER_STRIDE 16
__device__ __forceinline__ u64* counter_slot(u64* base) {
    // blocks that run at the same time have neighbouring indices: consecutive slots spread their atomics over 256 lines
    const unsigned b = blockIdx.x * 8u + blockIdx.y * 67u + blockIdx.z * 131u + (threadIdx.x >> 5);
    return base + (b & (DMF_COUNTER_SLOTS - 1u)) * DMF_COUNTER_STRIDE;
}

// ---- checked build (-DDMF_CHECKED; libdmf_b200_checked.so, tests/test_fuzz_gpu.py) ------------------------------------------
// compute-sanitizer is closed on the GPU pool this is developed on, so the kernels carry their own bounds checks: in the checked
// build every COMPUTED index into a grid or table is compared against its limit before use; a violation is counted in counter
// slot DMF_CNT_BOUNDS (and the index replaced by 0, so the run survives to report it).  The production build compiles the
// checks out: its marches rely on the proofs in DESIGN.md (padded index space, slab guards), which the fuzz test holds the
// checked build against.
#ifdef DMF_CHECKED
#define DMF_CHECK_IDX(idx, limit, counters) do { if ((unsigned long long)(idx) >= (unsigned long long)(limit)) { atomicAdd(counter_slot(counters) + 10, 1ull); (idx) = 0; } } while (0)
#else
#define DMF_CHECK_IDX(idx, limit, counters) do { } while (0)
#endif

// Thousands of rays of a view raise the same per-view flag.  A plain store per ray funnels ~10^4 same-address writes per
// view into one L2 slice, which serialises them (measured on B200: 1.07 of 2.9 ms per 128 views, and the SMs of the die
// that does not own the slice starve).  Read first: the line sits in L1, the flag only ever goes 0 -> 1, and a stale 0
// costs one redundant store at worst (the SM's own write-through store refreshes or drops its L1 copy).
__device__ __forceinline__ void raise_flag(int* p) {
    if (*p == 0) *p = 1;
}

// ---- fused exchange of a sharded sweep (dmf_comm.cuh) -------------------------------------------------------------------
// Candidate views are sharded over the GPUs of a box; every GPU needs every view's visibility row (the set-cover consumer).
// Rows are disjoint by view, so no collective is needed: the march kernel writes its views' rows in place in its OWN copy of
// the gathered [n_views][row_words] buffer, and the last block to finish a view (a ticket per view) pushes the finished row
// into every peer's copy through peer-mapped pointers -- plain 16-byte stores over NVLink, no cross-GPU atomics.  The block
// that completes the last view of the pass raises this GPU's sequence flag in every peer (system-scope release); consumers
// wait on their own copy of the flags (cuStreamWaitValue32), so the transfer overlaps the march view by view and nothing
// but the flag wait remains after the kernel.
#define DMF_MAX_PEERS 8
struct PubTable {
    unsigned* ticket;                    // [views of this pass] blocks that finished each view (zeroed per pass)
    unsigned* views_done;                // views of this pass whose row has been pushed (zeroed per pass)
    u64* peer_rows[DMF_MAX_PEERS];       // the gathered buffer (this pass's parity) in each OTHER member, as mapped here
    unsigned* peer_flag[DMF_MAX_PEERS];  // this rank's flag word in each other member's flag array
    const int* found_any;                // [views of this pass] per-view flag, copied into the row's extra word (null: 0)
    int n_peers;
    int n_views_pass;                    // views this GPU marches in the pass (all sub-launches together)
    int row0, row_step;                  // global row of local view j = row0 + j * row_step (views are dealt round-robin)
    unsigned row_words;                  // u64 words per gathered row: vis words + 1 (found flag), padded to an even count
    unsigned vis_words;                  // u64 words of visibility per row
    unsigned seq;                        // value to raise the flags to
    int enabled;
};

// Called by EVERY thread of the block at the very end of a march kernel (after the block's last visibility atomic).
// own_row: this view's row in this GPU's own gathered buffer (the kernel's visibility atomics went there).
__device__ __forceinline__ void publish_view_row(const PubTable& pub, u64* own_row, int view_local, unsigned blocks_per_view) {
    __shared__ int s_last;
    __syncthreads();                                       // every warp of the block has issued its visibility atomics
    if (threadIdx.x == 0) {
        __threadfence();                                   // cumulative: orders the block's atomics (observed through the barrier) before the ticket
        s_last = atomicAdd(pub.ticket + view_local, 1u) == blocks_per_view - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();                                       // the other blocks' atomics of this view are visible (they precede their tickets)
    const unsigned found = pub.found_any ? (unsigned)__ldcg(pub.found_any + view_local) : 0u;
    const size_t g_off = (size_t)(pub.row0 + view_local * pub.row_step) * pub.row_words;
    const unsigned n16 = pub.row_words >> 1;               // 16-byte pieces per row (row_words is even)
    if (threadIdx.x == 0) own_row[pub.vis_words] = (u64)found;
    for (unsigned i = threadIdx.x; i < n16; i += blockDim.x) {
        uint4 val = __ldcg(reinterpret_cast<const uint4*>(own_row) + i);     // L2: the L1 may hold a stale copy from the read-before-atomic test
        if (2u * i <= pub.vis_words && pub.vis_words < 2u * i + 2u) {          // the piece that holds the found word
            if (pub.vis_words & 1u) { val.z = found; val.w = 0u; } else { val.x = found; val.y = 0u; }
        }
        for (int p = 0; p < pub.n_peers; p++) reinterpret_cast<uint4*>(pub.peer_rows[p] + g_off)[i] = val;
    }
    __threadfence_system();                                // this thread's peer stores are performed before anything that follows
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(pub.views_done, 1u) == (unsigned)pub.n_views_pass - 1u) {
            __threadfence_system();                        // cumulative over the other finishing blocks' pushes (observed through views_done)
            for (int p = 0; p < pub.n_peers; p++) *reinterpret_cast<volatile unsigned*>(pub.peer_flag[p]) = pub.seq;
        }
    }
}

// Packed FP32 pairs (sm_100 FFMA2 / FADD2): two independent IEEE round-to-nearest operations per issued instruction,
// each half bit-identical to the scalar fmaf / __fadd_rn / __fsub_rn.  For loops that are bound by instruction issue,
// not by the FMA pipe.  A pair lives in one 64-bit register (lo = .x, hi = .y); pack/unpack are register-pair renames.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// HBM layout of a VoxelVolume (reference include/Volume.hpp:50-78):
//   bits     uint32 words of a linear bit grid over the PADDED index space [0,dim_x] x [0,dim_y] x [0,dim_z]
//            (pdim = dim + 1 per axis, z fastest like voxels_[x][y][z]); bit index = (x*pdim_y + y)*pdim_z + z.
//            The padding plane is always empty: a probe whose quotient rounds up to exactly `dim` (the reference
//            reads out of bounds there) lands on it, so the forward march needs no index range check.
//   prefix   uint32 per word: number of occupied voxels in all earlier words (rank directory; one load + one popc per hit)
//   rank2occ [n_occ] rank (linear order) -> index into occupied_cells_
//   bytes    optional per-voxel Chebyshev distance bytes over the same padded index space, 0 = occupied
//            (DMF_GRID_BYTE, dmf_distance.cuh)
//   macro    1 bit per 8x8x8-voxel macro cell: "contains an occupied voxel" (empty-space skipping)
//   noff/normals  CSR of Voxel::normals in occupied order
struct VolDev {
    const unsigned* __restrict__ bits;
    const unsigned* __restrict__ prefix;
    const unsigned* __restrict__ rank2occ;
    const unsigned char* __restrict__ bytes;
    const unsigned* __restrict__ macro;
    const unsigned* __restrict__ noff;
    const float* __restrict__ normals;
    const u64* __restrict__ occ_ids;
    int n_occ;
    unsigned n_cells; // pdim_x * pdim_y * pdim_z: entries of the padded grids
    int dim[3];      // xdim_, ydim_, zdim_
    int pdim[3];     // dim + 1
    int mdim[3];     // macro cells per axis = ceil(pdim / 8)
    double vmin[3];  // xmin_, ymin_, zmin_
    double delta[3]; // xdelta_, ydelta_, zdelta_
    double inv[3];   // RN(1/delta)
    double c0[3];    // RN(-vmin*inv)
    double half[3];  // delta/2.0
    double eps[3];   // |frac| below which the double fast quotient cannot be trusted (0 => axis is exact)
    float inv32[3];  // (float)inv
    float c32[3];    // (float)c0
    float err32[3];  // bound on |fmaf(p,inv32,c32) - reference quotient|  (0 => the float quotient is exact)
    float lo[3];     // largest float <= vmin   (validPoints: x<=xmin_  <=>  !(x > lo))
    float hi[3];     // smallest float >= vmax  (validPoints: x>=xmax_  <=>  !(x < hi))
    float rev_eps[3]; // bound (voxel units) on |reverse-march sample - its line point|; > 0.1 disables skipping there
    float rev_esafe;  // a line point at least this far from every face of its voxel shares the voxel with the reference's sample
    float ext[3];    // >= (vmax-vmin)/delta: the volume's extent in voxel units (dim <= ext < dim+1, constructVolume truncates)
    const unsigned* __restrict__ rev_perm;   // [n_occ] occupied ordinals in Morton order of their voxels: thread t of k_reverse<FAST> marches ordinal rev_perm[t] (null: t)
};

// The distance bytes hold the distance to the nearest OCCUPIED voxel only (dmf_distance.cuh).  A march that counts the samples it
// skips as in-bounds WITHOUT a slab interval of its own needs the voxel's distance to the volume's outermost voxel layer (index 0 or
// >= dim-1 on some axis; the L-inf distance to that set is the smallest per-axis distance) folded in: the result is the byte a
// transform with the outermost layer as a source would hold -- 0 = occupied, else max(1, min(d, border distance)).
__device__ __forceinline__ unsigned byte_with_border(const VolDev& v, unsigned d, int ix, int iy, int iz) {
    const int b = min(min(min(ix, iy), iz), min(min(v.dim[0] - 1 - ix, v.dim[1] - 1 - iy), v.dim[2] - 1 - iz));
    return d == 0u ? 0u : (unsigned)max(1, min((int)d, b));
}

// floor(((double)p - vmin) / delta) exactly as VoxelVolume::getVoxel (Volume.hpp:150-156) computes it, in double.
// q' = fma(p, 1/delta, -vmin/delta) differs from the reference quotient by < eps, so the floors agree unless q' is
// within eps of an integer; then (and only then) the reference's subtract + IEEE divide is executed.
// The floor itself uses the 1.5*2^52 shifter so no F2I/I2F conversions are needed.
__device__ __forceinline__ int voxel_index(float p, double vmin, double delta, double inv, double c0, double eps, unsigned& n_exact) {
    const double kShift = 6755399441055744.0;  // 1.5 * 2^52
    double q = fma((double)p, inv, c0);
    double s = __dadd_rn(q, kShift);
    int i = __double2loint(s);                 // rint(q)
    double f = __dsub_rn(q, __dsub_rn(s, kShift));   // q - rint(q), in [-0.5, 0.5]
    if (fabs(f) < eps) {
        double qe = __ddiv_rn(__dsub_rn((double)p, vmin), delta);
        n_exact++;
        return (int)floor(qe);
    }
    return i - (f < 0.0 ? 1 : 0);
}

// Float filter for the same index: q32 = fmaf(p, inv32, c32) is within err of the reference quotient (bound derived
// in DESIGN.md "voxel index filter", computed on the host per axis), so floor(q32) is the reference index whenever q32
// is at least err away from every integer.  Otherwise `unsafe` is raised and the caller redoes the probe in double.
__device__ __forceinline__ int voxel_index_f32(float p, float inv, float c, float err, bool& unsafe) {
    const float kShift = 12582912.0f;          // 1.5 * 2^23
    const float q = fmaf(p, inv, c);
    const float s = __fadd_rn(q, kShift);
    const float f = __fsub_rn(q, __fsub_rn(s, kShift));   // q - rint(q)
    unsafe = unsafe || (fabsf(f) < err);
    return (__float_as_int(s) - 0x4B400000) - (f < 0.0f ? 1 : 0);
}

__device__ __forceinline__ bool in_bounds(const VolDev& v, float x, float y, float z) {
    return x > v.lo[0] && x < v.hi[0] && y > v.lo[1] && y < v.hi[1] && z > v.lo[2] && z < v.hi[2];
}

__device__ __forceinline__ bool coords_valid(const VolDev& v, int x, int y, int z) {
    return (unsigned)x < (unsigned)v.dim[0] && (unsigned)y < (unsigned)v.dim[1] && (unsigned)z < (unsigned)v.dim[2];
}
// inside the padded index space (what the grids can be addressed with)
__device__ __forceinline__ bool coords_padded(const VolDev& v, int x, int y, int z) {
    return (unsigned)x < (unsigned)v.pdim[0] && (unsigned)y < (unsigned)v.pdim[1] && (unsigned)z < (unsigned)v.pdim[2];
}

__device__ __forceinline__ unsigned linear_index(const VolDev& v, int x, int y, int z) {   // < 2^31 for dims <= 1024
    return ((unsigned)x * (unsigned)v.pdim[1] + (unsigned)y) * (unsigned)v.pdim[2] + (unsigned)z;
}

// FMT 0: bit grid, 1: byte grid.  (x,y,z) must be inside the padded index space.
template <int FMT>
__device__ __forceinline__ bool occupied(const VolDev& v, int x, int y, int z) {
    const unsigned idx = linear_index(v, x, y, z);
    if (FMT == 0) return (__ldg(v.bits + (idx >> 5)) >> (idx & 31)) & 1u;
    return __ldg(v.bytes + idx) == 0;   // distance bytes: 0 = occupied (dmf_distance.cuh)
}

// index into occupied_cells_ of an occupied voxel
__device__ __forceinline__ int occupied_ordinal(const VolDev& v, int x, int y, int z) {
    const unsigned idx = linear_index(v, x, y, z);
    const unsigned w = idx >> 5;
    const unsigned rank = __ldg(v.prefix + w) + __popc(__ldg(v.bits + w) & ((1u << (idx & 31)) - 1u));
    return (int)__ldg(v.rank2occ + rank);
}

__device__ __forceinline__ u64 voxel_id(int x, int y, int z) {   // getHashId, Volume.hpp:143-148 (non-negative coords)
    return ((u64)(unsigned)x << 40) ^ (u64)(long long)(y << 20) ^ (u64)(long long)z;
}

// Eigen 3.3 Affine3f * Vector3f (rule E1 of oracle/dmf_oracle.hpp): ((m0*x + m1*y) + m2*z) + m3, no contraction.
__device__ __forceinline__ float affine_row(float m0, float m1, float m2, float m3, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, x), __fmul_rn(m1, y)), __fmul_rn(m2, z)), m3);
}
// Eigen 3.3 3-element reduction (rule E2): a0 + (a1 + a2)
__device__ __forceinline__ float sum3(float a0, float a1, float a2) { return __fadd_rn(a0, __fadd_rn(a1, a2)); }

// (camera_center - centroid).normalized() with centroid = point + delta/2 (RayTracingEngine.hpp:345-349), rule E3
__device__ __forceinline__ void view_direction(const VolDev& v, float px, float py, float pz, float camx, float camy, float camz,
                                               float& dx, float& dy, float& dz) {
    float cx = __double2float_rn(__dadd_rn((double)px, v.half[0]));
    float cy = __double2float_rn(__dadd_rn((double)py, v.half[1]));
    float cz = __double2float_rn(__dadd_rn((double)pz, v.half[2]));
    dx = __fsub_rn(camx, cx); dy = __fsub_rn(camy, cy); dz = __fsub_rn(camz, cz);
    float n2 = sum3(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz));
    if (n2 > 0.0f) {
        float s = __fsqrt_rn(n2);
        dx = __fdiv_rn(dx, s); dy = __fdiv_rn(dy, s); dz = __fdiv_rn(dz, s);
    }
}

// Parameters of the "angle in [0,90] degrees" test  degree(acos(n.v)) in [k_AngleMin,k_AngleMax]
// (RayTracingEngine.hpp:211-212, :362-364, :428-430; CommonUtilities.hpp:17).  The host bisects its own libm's
// acosf once (dmf_host.cuh): the test is true  <=>  dot_min <= d <= 1.  [band_lo, band_hi) is the range of d
// where the host libm was observed non-monotonic (empty if band_lo >= band_hi); hits inside it are counted as ties.
struct AngleTest { float dot_min, band_lo, band_hi; };

// some stored normal of occupied voxel `occ` passes the angle test against direction (dx,dy,dz)
__device__ __forceinline__ bool any_normal_faces(const VolDev& v, const AngleTest& at, int occ, float dx, float dy, float dz, unsigned& ties) {
    unsigned b = __ldg(v.noff + occ), e = __ldg(v.noff + occ + 1);
    for (unsigned j = b; j < e; j++) {
        float nx = __ldg(v.normals + 3 * (size_t)j), ny = __ldg(v.normals + 3 * (size_t)j + 1), nz = __ldg(v.normals + 3 * (size_t)j + 2);
        float d = sum3(__fmul_rn(nx, dx), __fmul_rn(ny, dy), __fmul_rn(nz, dz));
        if (d >= at.band_lo && d < at.band_hi) ties++;
        if (d >= at.dot_min && d <= 1.0f) return true;
    }
    return false;
}

// a / 1000.0f without the IEEE-division sequence (the reverse march divides three times per step:
// `centroid + v*double(depth)/1000.0`, RayTracingEngine.hpp:82,173).  r = RN(1/1000) and one Markstein correction.
// Exhaustion over all 2^32 float inputs on the B200 (k_selftest_div1000, tests/test_reverse_gpu.py::test_div1000_exhaustive)
// shows it equals __fdiv_rn(a, 1000.0f) bit for bit for every finite |a| > 2^-101; it differs only where the quotient is
// subnormal (|a| < 2^-122), for -0 and for +-inf.  Callers therefore use it only when |a| >= 2^-100 is guaranteed
// (k_reverse checks the direction components once per ray) and fall back to __fdiv_rn otherwise.
__device__ __forceinline__ float div1000_short(float a) {
    const float r = 1.0f / 1000.0f;
    const float q = __fmul_rn(a, r);
    return fmaf(fmaf(-1000.0f, q, a), r, q);
}
__device__ __forceinline__ float div1000(float a) {      // two corrections; same exactness domain, kept for the self-test
    const float r = 1.0f / 1000.0f;
    float q = __fmul_rn(a, r);
    q = fmaf(fmaf(-1000.0f, q, a), r, q);
    return fmaf(fmaf(-1000.0f, q, a), r, q);
}

// x86 cvttsd2si semantics (INT_MIN on NaN / overflow) for the reference's int(round(...)) casts (Camera.hpp:35-36)
__device__ __forceinline__ int to_int_x86(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return __double2int_rz(v);
}
