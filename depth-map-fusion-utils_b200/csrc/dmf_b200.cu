// dmf_b200.cu -- libdmf_b200.so: C ABI (include/dmf_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 (see build.py).
// There is no CPU fallback in this file: every compute entry point launches CUDA kernels or fails.
#include "dmf_host.cuh"
#include "dmf_forward.cuh"
#include "dmf_reverse.cuh"
#include "dmf_distance.cuh"
#include "dmf_integrate.cuh"
#include "dmf_volume.cuh"
#include "dmf_setcover.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <functional>
#include <new>

using namespace dmf;

namespace {

inline cudaStream_t pick_stream(dmf_ctx* c, void* s) { return s ? (cudaStream_t)s : c->stream; }
// All calls on a context share per-context scratch (view-start table, first-view array, inverse poses, projection tables, the
// timing events).  A call enqueued on a different stream than the previous one therefore first waits for that one's work.
inline int order_after_previous(dmf_ctx* c, cudaStream_t st) {
    if (c->last_stream_valid && c->last_stream != st) DMF_CUDA(cudaStreamWaitEvent(st, c->ev_last, 0));
    return 0;
}
inline int mark_last(dmf_ctx* c, cudaStream_t st) {
    DMF_CUDA(cudaEventRecord(c->ev_last, st));
    c->last_stream = st; c->last_stream_valid = true;
    return 0;
}
inline unsigned blocks_for(size_t n, unsigned threads, unsigned cap = 148 * 16) {
    size_t b = (n + threads - 1) / threads;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(b, cap));
}

inline unsigned long long fnv1a(const void* data, size_t n, unsigned long long h = 1469598103934665603ull) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// Run `enqueue` (which only enqueues work on `st`: no allocation, no synchronisation once its buffers are sized) directly, or --
// from the second call with the same key on -- as a captured CUDA graph.  See CallGraph.
template <class F>
int run_maybe_graphed(dmf_ctx* c, CallGraph& g, unsigned long long key, cudaStream_t st, F&& enqueue) {
    static const bool off = std::getenv("DMF_NO_GRAPH") != nullptr;
    if (off || g.disabled) return enqueue();
    const unsigned long long gen = alloc_generation().load();
    if (g.exec && g.key == key && g.gen == gen) {
        DMF_CUDA(cudaEventRecord(c->ev_h0, st));
        DMF_CUDA(cudaGraphLaunch(g.exec, st));
        DMF_CUDA(cudaEventRecord(c->ev_h1, st));
        c->hot_timed = true; c->launches += g.n_kernels;
        return 0;
    }
    if (g.cand_key == key && g.cand_gen == gen) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); g.disabled = true; return enqueue(); }
        const unsigned long long l0 = c->launches;
        c->capturing = true;
        const int rc = enqueue();
        c->capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (rc || e != cudaSuccess || !graph || alloc_generation().load() != gen) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            g.disabled = rc == 0 && alloc_generation().load() == gen;       // a capture that fails for no reason of ours: stop trying
            g.cand_key = 0;
            c->launches = l0;
            // nothing was executed by the aborted capture.  A failure INSIDE it may just be an operation that cannot be captured (a table
            // rebuild after an intervening call with other parameters synchronises the device): run the call plainly, it reports real errors
            return enqueue();
        }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.disabled = true; c->launches = l0; return enqueue(); }
        g.key = key; g.gen = gen; g.n_kernels = (unsigned)(c->launches - l0);
        c->launches = l0;
        DMF_CUDA(cudaEventRecord(c->ev_h0, st));
        DMF_CUDA(cudaGraphLaunch(g.exec, st));
        DMF_CUDA(cudaEventRecord(c->ev_h1, st));
        c->hot_timed = true; c->launches += g.n_kernels;
        return 0;
    }
    const int rc = enqueue();
    g.cand_key = key; g.cand_gen = alloc_generation().load();         // (this call may have sized buffers: remember the generation after it)
    return rc;
}

int fill_u32(dmf_ctx* c, cudaStream_t st, void* p, size_t n, unsigned val) {
    if (!n) return 0;
    k_fill_u32<<<blocks_for(n, 256), 256, 0, st>>>((unsigned*)p, n, val);
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    return 0;
}

// ---- device-wide exclusive scan of uint32, asynchronous -------------------------------------------------------
// d_out may alias d_in.  Blocks of SCAN_BLOCK elements, recursive on the block sums; the per-level sums live in the
// context's scan scratch.  If d_total is given it receives the grand total (device pointer).  Nothing synchronises.
int scan_u32_async(dmf_ctx* c, cudaStream_t st, const unsigned* d_in, unsigned* d_out, size_t n, unsigned* d_total) {
    if (n == 0) { if (d_total) DMF_CUDA(cudaMemsetAsync(d_total, 0, 4, st)); return 0; }
    std::vector<size_t> level_n;                      // elements per level: n, ceil(n/B), ... down to one block
    for (size_t m = n;; m = (m + SCAN_BLOCK - 1) / SCAN_BLOCK) { level_n.push_back(m); if (m <= (size_t)SCAN_BLOCK) break; }
    size_t scratch = 1;
    for (size_t l = 1; l < level_n.size(); l++) scratch += level_n[l];
    DMF_TRY(c->d_scan.reserve((scratch + 1) * 4));
    std::vector<unsigned*> lv(level_n.size() + 1);
    unsigned* sp = c->d_scan.as<unsigned>();
    for (size_t l = 1; l < level_n.size(); l++) { lv[l] = sp; sp += level_n[l]; }
    unsigned* top_total = sp;                           // one word: the sum of the last level
    for (size_t l = 0; l < level_n.size(); l++) {
        const unsigned* in = l == 0 ? d_in : lv[l];
        unsigned* out = l == 0 ? d_out : lv[l];
        unsigned* sums = l + 1 < level_n.size() ? lv[l + 1] : top_total;
        k_scan_block<<<(unsigned)((level_n[l] + SCAN_BLOCK - 1) / SCAN_BLOCK), SCAN_THREADS, 0, st>>>(in, out, level_n[l], sums);
        c->launches++;
    }
    for (size_t l = level_n.size() - 1; l-- > 0;) {
        unsigned* out = l == 0 ? d_out : lv[l];
        k_scan_add<<<(unsigned)((level_n[l] + SCAN_BLOCK - 1) / SCAN_BLOCK), SCAN_THREADS, 0, st>>>(out, level_n[l], lv[l + 1]);
        c->launches++;
    }
    if (d_total) DMF_CUDA(cudaMemcpyAsync(d_total, top_total, 4, cudaMemcpyDeviceToDevice, st));
    DMF_CUDA(cudaGetLastError());
    return 0;
}

// ---- volume: host-side constants ----------------------------------------------------------------------------------
// Validates (dim, bounds, delta) and fills the scalar half of VolDev: reciprocals, float thresholds, padded dims, the
// float-accumulated axes of the whole-grid loops.  No occupancy yet.
int set_volume_geometry(dmf_ctx* c, const double bounds[6], const double delta[3], const int dim[3]) {
    DMF_CUDA(cudaSetDevice(c->device));
    for (int a = 0; a < 3; a++) {
        if (dim[a] < 1 || dim[a] > 2048) return fail("volume dim %d out of range [1,2048] (voxel ids shift y by 20 bits as int, Volume.hpp:146)", dim[a]);
        if (!(delta[a] > 0) || !(bounds[2 * a + 1] > bounds[2 * a])) return fail("degenerate volume bounds/delta on axis %d", a);
        // the marches index the padded grid [0,dim] without a range check, which is sound only if every in-bounds sample's
        // quotient (p - vmin)/delta stays below dim + 1 (constructVolume's dim = int((max-min)/delta), Volume.hpp:121-123)
        const double extent = (bounds[2 * a + 1] - bounds[2 * a]) / delta[a];
        if (!(extent < (double)dim[a] + 1.0))
            return fail("volume axis %d: (max-min)/delta = %.17g does not fit dim %d (need < dim+1): dim, bounds and delta are inconsistent", a, extent, dim[a]);
    }
    if ((double)(dim[0] + 1) * (dim[1] + 1) * (dim[2] + 1) >= 4294967296.0) return fail("volume %dx%dx%d too large for the 32-bit linear voxel index", dim[0], dim[1], dim[2]);
    VolDev& v = c->vol;
    std::memcpy(c->bounds, bounds, sizeof c->bounds);
    c->voxel_size = delta[0] * delta[1] * delta[2];
    for (int a = 0; a < 3; a++) {
        const double vmin = bounds[2 * a], vmax = bounds[2 * a + 1];
        v.dim[a] = dim[a]; v.pdim[a] = dim[a] + 1; v.mdim[a] = (dim[a] + 1 + 7) / 8;
        v.vmin[a] = vmin; v.delta[a] = delta[a];
        v.inv[a] = 1.0 / delta[a];
        v.c0[a] = -vmin * v.inv[a];
        v.half[a] = delta[a] / 2.0;
        int e; const double mant = std::frexp(delta[a], &e);
        const bool pow2 = (mant == 0.5);
        if (pow2) v.eps[a] = 0.0;            // power-of-two delta: the double reciprocal multiply is exact
        else {
            double bound = std::ldexp((double)dim[a] + 2.0, -51) + std::ldexp(std::fabs(vmin) * v.inv[a], -53);
            v.eps[a] = std::max(4.0 * bound, std::ldexp(1.0, -30));
            if (v.eps[a] > 0.25) v.eps[a] = 1.0;    // hopeless conditioning: always take the exact path
        }
        // float filter: q32 = fmaf(p, inv32, c32).  |q32 - q_ref| <= 2^-24 * (|p|max/delta + |vmin|/delta + dim + 1) (+ double-level
        // terms); x2 safety.  Exact (err 0) when delta is a power of two and vmin == 0: then q32 == p/delta exactly.
        v.inv32[a] = (float)v.inv[a]; v.c32[a] = (float)v.c0[a];
        if (pow2 && vmin == 0.0) v.err32[a] = 0.0f;
        else {
            double pmax = std::max(std::fabs(vmin), std::fabs(vmax));
            double e32 = 2.0 * std::ldexp(pmax * v.inv[a] * 2.0 + std::fabs(vmin) * v.inv[a] * 2.0 + dim[a] + 2.0, -24) + v.eps[a];
            v.err32[a] = (float)std::min(e32 * (1.0 + 1e-6), 1.0);   // >= 0.5 means "always use the double path"
            if (!(e32 == e32)) v.err32[a] = 1.0f;
        }
        float lo = (float)vmin; if ((double)lo > vmin) lo = next_down(lo);
        float hi = (float)vmax; if ((double)hi < vmax) hi = next_up(hi);
        v.lo[a] = lo; v.hi[a] = hi;
        v.rev_eps[a] = (float)(std::ldexp(8.0, -24) * std::max(std::fabs(vmin), std::fabs(vmax)) * v.inv[a] * 2.0);
        v.ext[a] = next_up((float)((vmax - vmin) * v.inv[a] * (1.0 + 1e-7)));
    }
    {
        // 2 * rev_eps covers the reference sample's and the float line's distance from the ideal line; the dim-scaled term the
        // slope's rounding over a whole march (~4 ulp at the largest voxel coordinate); 2^-12 is plain slack
        const float re = std::max(v.rev_eps[0], std::max(v.rev_eps[1], v.rev_eps[2]));
        v.rev_esafe = 2.0f * re + (float)std::ldexp((double)std::max(dim[0], std::max(dim[1], dim[2])) + 2.0, -21) + (float)std::ldexp(1.0, -12);
    }
    // float-accumulated axes of the whole-grid loops (RayTracingEngine.hpp:54-56, :509-511)
    for (int a = 0; a < 3; a++) {
        std::vector<float> ax;
        double hi = bounds[2 * a + 1];
        for (float x = (float)bounds[2 * a]; x < hi; x = (float)(x + delta[a])) { ax.push_back(x); if (ax.size() > (1u << 22)) break; }
        c->n_axis[a] = (int)ax.size();
        DMF_TRY(c->d_axis[a].reserve(std::max<size_t>(ax.size(), 1) * 4));
        if (!ax.empty()) DMF_CUDA(cudaMemcpy(c->d_axis[a].p, ax.data(), ax.size() * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

// ---- volume: the march structures, built on the device -------------------------------------------------------------
// In: set_volume_geometry done; c->d_occ_ids (n_occ ids), c->d_noff (n_occ + 1) and c->d_normals (3 * n_normals floats)
// already hold the volume ON THE DEVICE (uploaded by the host, produced by K0, or received from a peer GPU).
// Out: bit grid, macro-cell bits + clearance, rank directory, rank -> ordinal table, centroid hashes, cleared marks.
// One synchronisation at the end (the error words).  Timed with CUDA events: dmf_volume_prepare_ms.
int build_volume_device(dmf_ctx* c, size_t n_occ, size_t n_normals) {
    VolDev& v = c->vol;
    cudaStream_t st = c->stream;
    const size_t nbits = (size_t)v.pdim[0] * v.pdim[1] * v.pdim[2];
    const size_t nwords = ((nbits + 31) / 32 + 7) / 8 * 8;           // whole 8-word (256-bit) rank blocks
    const size_t nmacro = (size_t)v.mdim[0] * v.mdim[1] * v.mdim[2];
    const size_t macro_words = (nmacro + 31) / 32;
    if (n_occ >= 0xFFFFFFFFull) return fail("too many occupied voxels");
    DMF_TRY(c->d_bricks.reserve(nwords * 4)); DMF_TRY(c->d_prefix.reserve(nwords * 4)); DMF_TRY(c->d_macro.reserve(macro_words * 4)); DMF_TRY(c->d_clearance.reserve(nmacro * 4));
    DMF_TRY(c->d_rank2occ.reserve(std::max<size_t>(n_occ, 1) * 4)); DMF_TRY(c->d_occ_ids.reserve(std::max<size_t>(n_occ, 1) * 8));
    DMF_TRY(c->d_noff.reserve((n_occ + 1) * 4)); DMF_TRY(c->d_normals.reserve(std::max<size_t>(3 * n_normals, 1) * 4));
    DMF_TRY(c->d_view_mark.reserve(std::max<size_t>(n_occ, 1) * 4)); DMF_TRY(c->d_first_view.reserve(std::max<size_t>(n_occ, 1) * 4));
    DMF_TRY(c->d_good_bits.reserve(((n_occ + 63) / 64 + 1) * 8));
    DMF_TRY(c->d_centroid_hash.reserve(std::max<size_t>(n_occ, 1) * 8));
    DMF_TRY(c->d_macro_dist[0].reserve(nmacro)); DMF_TRY(c->d_macro_dist[1].reserve(nmacro));
    DMF_TRY(c->d_err.reserve(16));
    v.bits = c->d_bricks.as<unsigned>(); v.prefix = c->d_prefix.as<unsigned>(); v.rank2occ = c->d_rank2occ.as<unsigned>(); v.macro = c->d_macro.as<unsigned>();
    v.noff = c->d_noff.as<unsigned>(); v.normals = c->d_normals.as<float>(); v.occ_ids = c->d_occ_ids.as<u64>();
    v.bytes = nullptr; v.n_occ = (int)n_occ; v.n_cells = (unsigned)nbits;
    c->n_occ = n_occ; c->n_normals = n_normals;
    c->bytes_built = false; c->auto_uses = 0; c->volume_epoch++;
    c->n_grid_words = nwords; c->observed_ready = false;     // a new volume starts unobserved

    DMF_CUDA(cudaEventRecord(c->ev_p0, st));
    DMF_CUDA(cudaMemsetAsync(c->d_bricks.p, 0, nwords * 4, st));
    DMF_CUDA(cudaMemsetAsync(c->d_macro.p, 0, macro_words * 4, st));
    DMF_CUDA(cudaMemsetAsync(c->d_err.p, 0xFF, 16, st));
    if (n_occ) {
        k_vol_mark<<<blocks_for(n_occ, 256), 256, 0, st>>>(v.occ_ids, (unsigned)n_occ, v, c->d_bricks.as<unsigned>(), c->d_macro.as<unsigned>(), c->d_err.as<unsigned>());
        k_vol_check_csr<<<blocks_for(n_occ, 256), 256, 0, st>>>(v.noff, (unsigned)n_occ, c->d_err.as<unsigned>());
        c->launches += 2;
    }
    k_vol_popc<<<blocks_for(nwords, 256), 256, 0, st>>>(v.bits, c->d_prefix.as<unsigned>(), nwords);
    c->launches++;
    DMF_TRY(scan_u32_async(c, st, c->d_prefix.as<unsigned>(), c->d_prefix.as<unsigned>(), nwords, nullptr));
    if (n_occ) {
        k_vol_rank2occ<<<blocks_for(n_occ, 256), 256, 0, st>>>(v.occ_ids, (unsigned)n_occ, v, v.bits, v.prefix, c->d_rank2occ.as<unsigned>());
        c->launches++;
    }
    // macro-cell clearance (k_forward_skip): Chebyshev distance, in cells, to the nearest blocked cell; cells outside the
    // grid are blocked (VIRTUAL_BORDER).  Three O(n) line passes over at most 257^3 bytes.
    {
        unsigned char *ma = c->d_macro_dist[0].as<unsigned char>(), *mb = c->d_macro_dist[1].as<unsigned char>();
        const int m0 = v.mdim[0], m1 = v.mdim[1], m2 = v.mdim[2];
        k_macro_seed<<<blocks_for(nmacro, 256), 256, 0, st>>>(v, v.macro, ma);
        k_dt_lines<2, true, false, EncodeNone><<<blocks_for((size_t)m0 * m1, 128), 128, 0, st>>>(ma, mb, m0, m1, m2, EncodeNone());
        k_dt_lines<1, true, false, EncodeNone><<<blocks_for((size_t)m0 * m2, 128), 128, 0, st>>>(mb, ma, m0, m1, m2, EncodeNone());
        k_dt_lines<0, true, false, EncodeNone><<<blocks_for((size_t)m1 * m2, 128), 128, 0, st>>>(ma, mb, m0, m1, m2, EncodeNone());
        k_macro_clearance<<<blocks_for(nmacro, 256), 256, 0, st>>>(mb, c->d_clearance.as<float>(), (unsigned)nmacro);
        c->launches += 5;
    }
    DMF_CUDA(cudaMemsetAsync(c->d_view_mark.p, 0, std::max<size_t>(n_occ, 1) * 4, st));
    DMF_CUDA(cudaMemsetAsync(c->d_good_bits.p, 0, ((n_occ + 63) / 64 + 1) * 8, st));
    DMF_TRY(fill_u32(c, st, c->d_first_view.p, std::max<size_t>(n_occ, 1), 0x7fffffffu));
    // centroid hashes of the occupied voxels (RayTracingEngine.hpp:151-163), view independent
    if (n_occ) {
        k_centroid_hash<<<blocks_for(n_occ, 256), 256, 0, st>>>(c->vol, c->d_centroid_hash.as<u64>());
        c->launches++;
    }
    // work order of the fast reverse march: ordinals sorted by the Morton code of their voxels (dmf_volume.cuh)
    v.rev_perm = nullptr;
    if (n_occ >= 64) {
        // (the context's call scratch serves as temporaries: grow-only, stream-ordered, nothing to free or wait for)
        DevBuf &keys_in = c->d_tmp_a, &keys_out = c->d_tmp_b, &vals_in = c->d_misc[0], &tmp = c->d_misc[1];
        DMF_TRY(keys_in.reserve(n_occ * 8)); DMF_TRY(keys_out.reserve(n_occ * 8)); DMF_TRY(vals_in.reserve(n_occ * 4)); DMF_TRY(c->d_rev_perm.reserve(n_occ * 4));
        k_morton_keys<<<blocks_for(n_occ, 256), 256, 0, st>>>(v.occ_ids, (unsigned)n_occ, keys_in.as<u64>(), vals_in.as<unsigned>());
        int axis_bits = 1;                                                            // the keys interleave three coordinates < dim: sort only the bits they use
        while (axis_bits < 21 && (1 << axis_bits) < std::max(v.dim[0], std::max(v.dim[1], v.dim[2]))) axis_bits++;
        const int end_bit = 3 * axis_bits;
        size_t tmp_bytes = 0;
        DMF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.as<u64>(), keys_out.as<u64>(), vals_in.as<unsigned>(), c->d_rev_perm.as<unsigned>(), (int)n_occ, 0, end_bit, st));
        DMF_TRY(tmp.reserve(std::max<size_t>(tmp_bytes, 16)));
        DMF_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_in.as<u64>(), keys_out.as<u64>(), vals_in.as<unsigned>(), c->d_rev_perm.as<unsigned>(), (int)n_occ, 0, end_bit, st));
        c->launches += 2;
        if (!std::getenv("DMF_REVERSE_NO_PERM")) v.rev_perm = c->d_rev_perm.as<unsigned>();
    }
    DMF_CUDA(cudaEventRecord(c->ev_p1, st));
    DMF_CUDA(cudaGetLastError());
    unsigned err[4];
    DMF_CUDA(cudaMemcpyAsync(err, c->d_err.p, 16, cudaMemcpyDeviceToHost, st));
    DMF_CUDA(cudaStreamSynchronize(st));
    c->prepare_timed = true;
    if (err[0] != 0xFFFFFFFFu) { c->vol_set = false; return fail("occupied id #%u lies outside the %dx%dx%d grid", err[0], v.dim[0], v.dim[1], v.dim[2]); }
    if (err[1] != 0xFFFFFFFFu) { c->vol_set = false; return fail("duplicate occupied id (#%u)", err[1]); }
    if (err[2] != 0xFFFFFFFFu) { c->vol_set = false; return fail("normal_offsets not monotone at voxel %u", err[2]); }
    c->vol_set = true;
    return 0;
}

// ---- volume upload from host arrays -----------------------------------------------------------------------------------
int upload_volume(dmf_ctx* c, const double bounds[6], const double delta[3], const int dim[3],
                  const uint64_t* ids, size_t n_occ, const uint32_t* noff, const float* normals) {
    if (n_occ && !ids) return fail("null occupied id list");
    if (noff) {
        // CSR of the per-voxel normal lists: any_normal_faces walks it on the device without further checks
        if (noff[0] != 0) return fail("normal_offsets[0] must be 0 (got %u)", noff[0]);
        for (size_t i = 0; i < n_occ; i++) if (noff[i + 1] < noff[i]) return fail("normal_offsets not monotone at voxel %zu (%u > %u)", i, noff[i], noff[i + 1]);
        if (noff[n_occ] && !normals) return fail("normal_offsets describe %u normals but the normals array is null", noff[n_occ]);
    }
    DMF_TRY(set_volume_geometry(c, bounds, delta, dim));
    c->vol_set = false;
    // host mirrors (dmf_volume_get_*): the caller's arrays as they are
    c->h_occ.assign(ids, ids + n_occ);
    c->h_noff.assign(n_occ + 1, 0);
    if (noff) c->h_noff.assign(noff, noff + n_occ + 1);
    const size_t n_normals = c->h_noff[n_occ];
    c->h_normals.assign(3 * n_normals, 0.f);
    if (normals && n_normals) c->h_normals.assign(normals, normals + 3 * n_normals);
    c->mirror_valid = true;
    DMF_TRY(c->d_occ_ids.reserve(std::max<size_t>(n_occ, 1) * 8)); DMF_TRY(c->d_noff.reserve((n_occ + 1) * 4)); DMF_TRY(c->d_normals.reserve(std::max<size_t>(3 * n_normals, 1) * 4));
    cudaStream_t st = c->stream;
    if (n_occ) DMF_CUDA(cudaMemcpyAsync(c->d_occ_ids.p, c->h_occ.data(), n_occ * 8, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_noff.p, c->h_noff.data(), (n_occ + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_normals) DMF_CUDA(cudaMemcpyAsync(c->d_normals.p, c->h_normals.data(), 3 * n_normals * 4, cudaMemcpyHostToDevice, st));
    return build_volume_device(c, n_occ, n_normals);
}

// host mirrors of a volume that was built or received on the device: fetched on first use
int ensure_host_mirror(dmf_ctx* c) {
    if (c->mirror_valid) return 0;
    DMF_CUDA(cudaSetDevice(c->device));
    c->h_occ.resize(c->n_occ); c->h_noff.assign(c->n_occ + 1, 0); c->h_normals.resize(3 * c->n_normals);
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    if (c->n_occ) DMF_CUDA(cudaMemcpy(c->h_occ.data(), c->d_occ_ids.p, c->n_occ * 8, cudaMemcpyDeviceToHost));
    DMF_CUDA(cudaMemcpy(c->h_noff.data(), c->d_noff.p, (c->n_occ + 1) * 4, cudaMemcpyDeviceToHost));
    if (c->n_normals) DMF_CUDA(cudaMemcpy(c->h_normals.data(), c->d_normals.p, 3 * c->n_normals * 4, cudaMemcpyDeviceToHost));
    c->mirror_valid = true;
    return 0;
}

// carve mode: the observed-voxel bit grid, allocated and zeroed on first use (and again after every volume upload)
int ensure_observed(dmf_ctx* c, cudaStream_t st) {
    if (c->observed_ready) return 0;
    DMF_TRY(c->d_observed.reserve(std::max<size_t>(c->n_grid_words, 8) * 4));
    DMF_CUDA(cudaMemsetAsync(c->d_observed.p, 0, std::max<size_t>(c->n_grid_words, 8) * 4, st));
    c->observed_ready = true;
    return 0;
}

__global__ void k_observed_counts(const unsigned* __restrict__ obs, const unsigned* __restrict__ occ, size_t n_words, u64* out) {
    unsigned long long n_obs = 0, n_hit = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned o = obs[i];
        n_obs += (unsigned)__popc(o); n_hit += (unsigned)__popc(o & __ldg(occ + i));
    }
    for (int s = 16; s > 0; s >>= 1) { n_obs += __shfl_down_sync(0xffffffffu, n_obs, s); n_hit += __shfl_down_sync(0xffffffffu, n_hit, s); }
    if ((threadIdx.x & 31) == 0) { if (n_obs) atomicAdd(out, n_obs); if (n_hit) atomicAdd(out + 1, n_hit); }
}

// DMF_GRID_BYTE: the per-voxel Chebyshev distance bytes (dmf_distance.cuh), built on first use
int ensure_bytes(dmf_ctx* c, cudaStream_t st) {
    if (c->bytes_built) return 0;
    VolDev& v = c->vol;
    const size_t n = (size_t)v.pdim[0] * v.pdim[1] * v.pdim[2];
    DMF_TRY(c->d_bytes.reserve(n));
    DMF_TRY(c->d_dt_tmp.reserve(n));
    unsigned char *bytes = c->d_bytes.as<unsigned char>(), *tmp = c->d_dt_tmp.as<unsigned char>();
    const unsigned nlines = (unsigned)v.pdim[0] * (unsigned)v.pdim[1];
    DMF_CUDA(cudaEventRecord(c->ev_b0, st));
    k_dt_z<<<blocks_for(nlines, 8, 148 * 32), 256, 0, st>>>(v, bytes);
    k_dt_lines<1, false, false, EncodeNone><<<blocks_for((size_t)v.pdim[0] * v.pdim[2], 128, 148 * 32), 128, 0, st>>>(bytes, tmp, v.pdim[0], v.pdim[1], v.pdim[2], EncodeNone());
    k_dt_lines<0, false, true, EncodeNone><<<blocks_for((size_t)v.pdim[1] * v.pdim[2], 128, 148 * 32), 128, 0, st>>>(tmp, bytes, v.pdim[0], v.pdim[1], v.pdim[2], EncodeNone());
    DMF_CUDA(cudaEventRecord(c->ev_b1, st));
    c->launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("distance transform failed: %s", cudaGetErrorString(e));
    v.bytes = bytes;
    c->bytes_built = true; c->bytes_timed = true;
    return 0;
}

int ensure_tables(dmf_ctx* c, int z0, int zdelta, int cstride, int rstride, cudaStream_t st) {
    TableKey k; std::memset(&k, 0, sizeof k);
    std::memcpy(k.K, c->K, sizeof k.K); k.H = c->H; k.W = c->W; k.z0 = z0; k.zdelta = zdelta; k.cstride = cstride; k.rstride = rstride;
    if (c->tables_valid && k == c->tkey) return 0;
    // for(z_depth=z0; z_depth<k_ZMax*1000; z_depth+=zdelta)   (RayTracingEngine.hpp:239,280)
    int S = (1000 - z0 + zdelta - 1) / zdelta;
    int Wc = (c->W + cstride - 1) / cstride, Hc = (c->H + rstride - 1) / rstride;
    DMF_TRY(c->d_xtab.reserve((size_t)S * Wc * 4)); DMF_TRY(c->d_ytab.reserve((size_t)S * Hc * 4)); DMF_TRY(c->d_ztab.reserve((size_t)S * 4));
    DMF_TRY(c->d_dcx.reserve((size_t)Wc * 4)); DMF_TRY(c->d_dcy.reserve((size_t)Hc * 4));
    // the tables may still be in use by work queued on another stream of this context
    DMF_CUDA(cudaDeviceSynchronize());
    k_build_tables<<<blocks_for((size_t)S * (Wc + Hc + 1), 256), 256, 0, st>>>(c->d_xtab.as<float>(), c->d_ytab.as<float>(), c->d_ztab.as<float>(), c->d_dcx.as<float>(), c->d_dcy.as<float>(),
        S, Wc, Hc, cstride, rstride, z0, zdelta, (double)c->K[0], (double)c->K[2], (double)c->K[4], (double)c->K[5]);
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    {
        const double fx = c->K[0], cx = c->K[2], fy = c->K[4], cy = c->K[5];
        c->dcx_max = (float)(std::max(std::fabs(0.0 - cx), std::fabs((double)(c->W - 1) - cx)) / std::fabs(fx) * 1.0001);
        c->dcy_max = (float)(std::max(std::fabs(0.0 - cy), std::fabs((double)(c->H - 1) - cy)) / std::fabs(fy) * 1.0001);
    }
    c->S = S; c->Wc = Wc; c->Hc = Hc; c->tkey = k; c->tables_valid = true;
    return 0;
}

template <int MODE>
void launch_forward_fmt(const FwdArgs& a, int fmt, bool skip, bool two_probe, dim3 grid, cudaStream_t st) {
    if (a.observed) {
        // carve mode: every sample has to be located, so none can be skipped; with the distance bytes the line-first kernel
        // finds the hit and then locates the samples on the line (carve_on_line), otherwise the brute-force march records them
        if constexpr (MODE != 4) {
            if (skip && fmt == DMF_GRID_BYTE && !two_probe) {
                dim3 g((a.Wc + SKIP_TILE_W - 1) / SKIP_TILE_W, (a.Hc + SKIP_TILE_H - 1) / SKIP_TILE_H, grid.z);
                const bool exact = a.vol.err32[0] == 0.0f && a.vol.err32[1] == 0.0f && a.vol.err32[2] == 0.0f;
                if (exact) k_forward_line<MODE, true, true><<<g, SKIP_THREADS, 0, st>>>(a);
                else k_forward_line<MODE, false, true><<<g, SKIP_THREADS, 0, st>>>(a);
            } else if (fmt == DMF_GRID_BYTE) k_forward<MODE, 1, true><<<grid, FWD_THREADS, 0, st>>>(a);
            else k_forward<MODE, 0, true><<<grid, FWD_THREADS, 0, st>>>(a);
        }
    } else if (skip) {
        dim3 g((a.Wc + SKIP_TILE_W - 1) / SKIP_TILE_W, (a.Hc + SKIP_TILE_H - 1) / SKIP_TILE_H, grid.z);
        const bool exact = a.vol.err32[0] == 0.0f && a.vol.err32[1] == 0.0f && a.vol.err32[2] == 0.0f;
        if (fmt == DMF_GRID_BYTE) {
            if (two_probe) {
                if (exact) k_forward_dist<MODE, true><<<g, SKIP_THREADS, 0, st>>>(a);
                else k_forward_dist<MODE, false><<<g, SKIP_THREADS, 0, st>>>(a);
            } else {
                if (exact) k_forward_line<MODE, true, false><<<g, SKIP_THREADS, 0, st>>>(a);
                else k_forward_line<MODE, false, false><<<g, SKIP_THREADS, 0, st>>>(a);
            }
        } else k_forward_skip<MODE, 0><<<g, SKIP_THREADS, 0, st>>>(a);
    } else {
        if (fmt == DMF_GRID_BYTE) k_forward<MODE, 1, false><<<grid, FWD_THREADS, 0, st>>>(a);
        else k_forward<MODE, 0, false><<<grid, FWD_THREADS, 0, st>>>(a);
    }
}

struct FwdPlan { int z0, cstride, rstride, grid_format; };
int plan_forward(dmf_ctx* c, const dmf_forward_params* p, FwdPlan& pl) {
    if (!c->cam_set) return fail("dmf_set_camera has not been called");
    if (!c->vol_set) return fail("no volume uploaded");
    if (p->mode < 0 || p->mode > 4) return fail("bad mode %d", p->mode);
    if (p->zdelta < 1) return fail("zdelta must be >= 1 (the reference loops forever on zdelta <= 0)");
    if (p->grid_format != DMF_GRID_BIT && p->grid_format != DMF_GRID_BYTE && p->grid_format != DMF_GRID_AUTO) return fail("bad grid_format %d", p->grid_format);
    // DMF_GRID_AUTO: the bit grid is ready as soon as the volume is uploaded; the distance bytes cost a one-off transform (2.6 ms at
    // 512^3, 16 ms at 1024^3) and make every later march ~4x faster.  Use them once they exist (the reverse march builds them) or
    // from the second forward call on one volume.
    pl.grid_format = p->grid_format;
    if (p->grid_format == DMF_GRID_AUTO) pl.grid_format = (c->bytes_built || ++c->auto_uses >= 2) ? DMF_GRID_BYTE : DMF_GRID_BIT;
    if ((p->flags & DMF_FWD_CARVE) && p->mode == DMF_MODE_MINIMUM)
        return fail("DMF_FWD_CARVE is not defined for MINIMUM mode (rayTraceAndGetMinimum returns mid-plane: the samples it visits depend on the pixel order)");
    pl.z0 = p->mode == DMF_MODE_MINIMUM ? 5 : 10;                                        // :239 vs :280,:327,:396,:461
    pl.cstride = pl.rstride = p->sparse ? (p->mode == DMF_MODE_MINIMUM ? 10 : 5) : 1;    // :236-237 vs :277-278
    return 0;
}

// Enqueue the march for n_views poses already on the device.  out holds device pointers.
// ids bookkeeping (first_key / ray_key / ray_occ) is passed separately; all three null if ids are not wanted.
// sub_views > 0 marches the batch as several launches of that many views (same buffers, same results) and calls
// after_sub(first_view, n, stream it was launched on) once each is enqueued, so that a caller can start copying finished
// views while later ones run.  The launches alternate between `st` and the context's auxiliary stream: they are
// independent, and this way the first blocks of one launch fill the SMs that the tail of the previous one leaves idle
// (measured: ~0.1 ms per launch boundary otherwise).  `st` is joined with the auxiliary stream before returning.
int enqueue_forward(dmf_ctx* c, const dmf_forward_params* p, const FwdPlan& pl, const float* d_poses, int n_views, int view_id0,
                    const dmf_forward_out& out, unsigned* first_key, unsigned* ray_key, int* ray_occ, cudaStream_t st,
                    int sub_views = 0, const std::function<int(int, int, cudaStream_t)>* after_sub = nullptr,
                    unsigned vis_stride32 = 0, const PubTable* pub = nullptr) {
    if (n_views <= 0) return 0;
    if (n_views > 65535) return fail("at most 65535 views per launch (got %d)", n_views);
    DMF_TRY(ensure_tables(c, pl.z0, p->zdelta, pl.cstride, pl.rstride, st));
    if (pl.grid_format == DMF_GRID_BYTE) DMF_TRY(ensure_bytes(c, st));
    const bool carve = (p->flags & DMF_FWD_CARVE) != 0;
    if (carve) DMF_TRY(ensure_observed(c, st));
    const size_t HW = (size_t)c->H * c->W;
    const size_t vis_words64 = (c->n_occ + 63) / 64;
    const bool sparse_lattice = pl.cstride > 1;
    if (out.depth_mm && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.depth_mm, 0xFF, n_views * HW * 4, st));
    if (out.depth_u16 && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.depth_u16, 0xFF, n_views * HW * 2, st));
    if (out.hit_voxel && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.hit_voxel, 0xFF, n_views * HW * 8, st));
    if (out.points && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.points, 0, n_views * HW * 12, st));
    // (a sharded sweep passes its own row pitch: the rows live interleaved in the gathered buffer and the caller zeroed them)
    if (out.visibility && vis_words64 && !vis_stride32) DMF_CUDA(cudaMemsetAsync(out.visibility, 0, (size_t)n_views * vis_words64 * 8, st));
    if (out.found_any) DMF_CUDA(cudaMemsetAsync(out.found_any, 0, (size_t)n_views * 4, st));
    if (p->mode == DMF_MODE_MINIMUM) {
        if (!out.min_depth) return fail("MINIMUM mode needs out.min_depth");
        DMF_TRY(fill_u32(c, st, out.min_depth, n_views, 0x7fffffffu));
    }
    if (first_key) DMF_TRY(fill_u32(c, st, first_key, (size_t)n_views * c->n_occ, 0xFFFFFFFFu));

    // the kernels index pixels, lattice rays and visibility words with 32 bits
    if ((double)n_views * c->H * c->W >= 4294967296.0 || (double)n_views * (double)vis_words64 * 2.0 >= 4294967296.0)
        return fail("%d views of %dx%d exceed the 32-bit pixel index of one launch: split the batch", n_views, c->W, c->H);
    FwdArgs a;
    a.vol = c->vol; a.angle = c->angle; a.poses = d_poses;
    a.xtab = c->d_xtab.as<float>(); a.ytab = c->d_ytab.as<float>(); a.ztab = c->d_ztab.as<float>();
    a.S = c->S; a.Wc = c->Wc; a.Hc = c->Hc; a.W = c->W; a.H = c->H; a.cstride = pl.cstride; a.rstride = pl.rstride; a.z0 = pl.z0; a.zdelta = p->zdelta;
    a.depth = out.depth_mm; a.depth16 = out.depth_u16; a.points = out.points; a.hit_voxel = (u64*)out.hit_voxel;
    a.vis = (unsigned*)out.visibility; a.vis_words32 = (int)(vis_words64 * 2);
    a.vis_stride32 = vis_stride32 ? vis_stride32 : (unsigned)(vis_words64 * 2);
    if (pub) a.pub = *pub; else std::memset(&a.pub, 0, sizeof a.pub);
    a.found_any = out.found_any; a.min_depth = out.min_depth;
    a.first_key = first_key; a.ray_key = ray_key; a.ray_occ = ray_occ;
    a.first_view = c->d_first_view.as<int>(); a.good_bits = c->d_good_bits.as<unsigned>(); a.view_mark = c->d_view_mark.as<int>();
    a.counters = (p->flags & DMF_FWD_NO_COUNTERS) ? nullptr : c->d_counters.as<u64>();
    a.observed = carve ? c->d_observed.as<unsigned>() : nullptr;
    a.dcx = c->d_dcx.as<float>(); a.dcy = c->d_dcy.as<float>(); a.clearance = c->d_clearance.as<float>();
    a.dcx_max = c->dcx_max; a.dcy_max = c->dcy_max;
    DMF_TRY(c->d_kstart.reserve((size_t)n_views * 8 + 64 + (size_t)n_views * 64));
    a.kstart = c->d_kstart.as<int>(); a.veps = c->d_kstart.as<float>() + n_views;
    a.viewrec = (float4*)(((uintptr_t)(c->d_kstart.as<char>() + (size_t)n_views * 8) + 63) & ~(uintptr_t)63);       // 64-byte records behind the two arrays
    a.z0m = (float)pl.z0 * 0.001f; a.zdm = (float)p->zdelta * 0.001f; a.Sf = (float)c->S;
    a.pnyz = (unsigned)c->vol.pdim[1] * (unsigned)c->vol.pdim[2];
    a.bias = 0x4B400000u * (a.pnyz + (unsigned)c->vol.pdim[2] + 1u);
    a.last = a.pnyz * (unsigned)c->vol.pdim[0] - 1u;
    const bool skip = !(p->flags & DMF_FWD_NO_SKIP);
    const bool byte_skip = skip && pl.grid_format == DMF_GRID_BYTE;
    const bool two_probe = (p->flags & DMF_FWD_TWO_PROBE) != 0;
    if (sub_views <= 0 || sub_views > n_views) sub_views = n_views;
    if (!c->capturing) DMF_CUDA(cudaEventRecord(c->ev_h0, st));
    a.tile_rec = nullptr; a.tiles_x = (c->Wc + TILE_RAYS - 1) / TILE_RAYS; a.tiles_per_view = a.tiles_x * ((c->Hc + TILE_RAYS - 1) / TILE_RAYS);
    if (byte_skip) { a.view0 = 0; k_view_start<<<(n_views + 127) / 128, 128, 0, st>>>(a, n_views, const_cast<int*>(a.kstart)); c->launches++; }
    if (byte_skip && !two_probe) {                                           // the line-first kernel runs: per-tile sample intervals + cone pre-march
        if (c->S > 4095) return fail("internal: %d samples per ray exceed the 12-bit fields of the tile records", c->S);
        DMF_TRY(c->d_tile_rec.reserve((size_t)n_views * a.tiles_per_view * 8));
        k_tile_start<<<dim3((a.tiles_per_view + 127) / 128, n_views), 128, 0, st>>>(a, c->d_tile_rec.as<u64>());
        a.tile_rec = c->d_tile_rec.as<u64>();
        c->launches++;
    }
    const bool split = sub_views < n_views;
    if (split) { DMF_CUDA(cudaEventRecord(c->ev_fork, st)); DMF_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0)); }
    int n_launch = 0;
    for (int v0 = 0; v0 < n_views; v0 += sub_views, n_launch++) {
        const int nv = std::min(sub_views, n_views - v0);
        cudaStream_t ls = (n_launch & 1) ? c->aux_stream : st;
        a.view0 = v0;
        dim3 grid((c->Wc + FWD_TILE_W - 1) / FWD_TILE_W, (c->Hc + FWD_TILE_H - 1) / FWD_TILE_H, nv);
        switch (p->mode) {
            case 0: launch_forward_fmt<0>(a, pl.grid_format, skip, two_probe, grid, ls); break;
            case 1: launch_forward_fmt<1>(a, pl.grid_format, skip, two_probe, grid, ls); break;
            case 2: launch_forward_fmt<2>(a, pl.grid_format, skip, two_probe, grid, ls); break;
            case 3: launch_forward_fmt<3>(a, pl.grid_format, skip, two_probe, grid, ls); break;
            default: launch_forward_fmt<4>(a, pl.grid_format, skip, two_probe, grid, ls); break;
        }
        c->launches++;
        if (after_sub) DMF_TRY((*after_sub)(v0, nv, ls));
    }
    if (split) { DMF_CUDA(cudaEventRecord(c->ev_join, c->aux_stream)); DMF_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0)); }
    if (!c->capturing) { DMF_CUDA(cudaEventRecord(c->ev_h1, st)); c->hot_timed = true; }
    DMF_CUDA(cudaGetLastError());
    if (p->mode == DMF_MODE_CLASSIFY && c->n_occ && !c->defer_first_view) {
        k_apply_first_view<<<blocks_for(c->n_occ, 256, 1u << 30), 256, 0, st>>>(c->d_view_mark.as<int>(), c->d_first_view.as<int>(), (int)c->n_occ, view_id0);
        c->launches++;
        DMF_CUDA(cudaGetLastError());
    }
    if (p->mode == DMF_MODE_MINIMUM) {
        k_finish_min_depth<<<blocks_for(n_views, 256, 1u << 30), 256, 0, st>>>(out.min_depth, n_views);
        c->launches++;
        DMF_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace

// ======================================================= C ABI ==============================================
extern "C" {

int dmf_version(void) {                                  // 2xx: round 2 ABI; odd: the checked build (-DDMF_CHECKED)
#ifdef DMF_CHECKED
    return 201;
#else
    return 200;
#endif
}

int dmf_host_angle_test(float out[3]) {
    if (!out) return fail("null argument");
    AngleTest t = bisect_angle_test();
    out[0] = t.dot_min; out[1] = t.band_lo; out[2] = t.band_hi;
    return 0;
}
const char* dmf_last_error(void) { return last_error().c_str(); }

int dmf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void* dmf_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void dmf_host_free(void* p) { if (p) cudaFreeHost(p); }

int dmf_create(dmf_ctx** out, int device) {
    if (!out) return fail("dmf_create: null out pointer");
    *out = nullptr;
    int n = dmf_device_count();
    if (n <= 0) return fail("no CUDA device visible: libdmf_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail("device %d out of range (have %d)", device, n);
    DMF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DMF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("device %d is sm_%d%d; libdmf_b200 carries sm_100a code only", device, prop.major, prop.minor);
    dmf_ctx* c = new (std::nothrow) dmf_ctx();
    if (!c) return fail("out of host memory");
    c->device = device;
    DMF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    DMF_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    DMF_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    DMF_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); DMF_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    DMF_CUDA(cudaEventCreate(&c->ev_k0)); DMF_CUDA(cudaEventCreate(&c->ev_k1));
    DMF_CUDA(cudaEventCreate(&c->ev_h0)); DMF_CUDA(cudaEventCreate(&c->ev_h1));
    DMF_CUDA(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
    DMF_CUDA(cudaEventCreate(&c->ev_p0)); DMF_CUDA(cudaEventCreate(&c->ev_p1)); DMF_CUDA(cudaEventCreate(&c->ev_b0)); DMF_CUDA(cudaEventCreate(&c->ev_b1));
    for (int i = 0; i < 2; i++) {
        DMF_CUDA(cudaEventCreateWithFlags(&c->ev_compute[i], cudaEventDisableTiming));
        DMF_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
    DMF_TRY(c->d_counters.reserve(DMF_COUNTER_SLOTS * DMF_COUNTER_STRIDE * 8));
    DMF_CUDA(cudaMemset(c->d_counters.p, 0, DMF_COUNTER_SLOTS * DMF_COUNTER_STRIDE * 8));
    c->angle = bisect_angle_test();
    *out = c;
    return 0;
}

void dmf_destroy(dmf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&c->d_bricks, &c->d_macro, &c->d_clearance, &c->d_dcx, &c->d_dcy, &c->d_prefix, &c->d_rank2occ, &c->d_bytes, &c->d_noff, &c->d_normals, &c->d_occ_ids, &c->d_centroid_hash,
                      &c->d_view_mark, &c->d_good_bits, &c->d_first_view, &c->d_observed, &c->d_axis[0], &c->d_axis[1], &c->d_axis[2], &c->d_xtab, &c->d_ytab, &c->d_ztab, &c->d_kstart,
                      &c->d_poses[0], &c->d_poses[1], &c->d_inv_poses, &c->d_first_key, &c->d_ray_key, &c->d_ray_occ, &c->d_tmp_a, &c->d_tmp_b,
                      &c->d_out_occ, &c->d_n_ids, &c->d_offsets, &c->d_ids, &c->d_misc[0], &c->d_misc[1], &c->d_misc[2], &c->d_misc[3], &c->d_counters,
                      &c->d_scan, &c->d_dt_tmp, &c->d_macro_dist[0], &c->d_macro_dist[1], &c->d_err, &c->d_tile_rec, &c->d_rev_perm};
    c->graph_fwd_ids.drop(); c->graph_rev_ids.drop();
    for (auto* b : bufs) b->release();
    c->stage.release();
    for (int i = 0; i < 2; i++) for (int j = 0; j < 8; j++) c->d_out[i][j].release();
    for (int i = 0; i < 2; i++) { if (c->ev_compute[i]) cudaEventDestroy(c->ev_compute[i]); if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); }
    for (cudaEvent_t e : {c->ev_p0, c->ev_p1, c->ev_b0, c->ev_b1, c->ev_last}) if (e) cudaEventDestroy(e);
    if (c->ev_h0) cudaEventDestroy(c->ev_h0);
    if (c->ev_h1) cudaEventDestroy(c->ev_h1);
    if (c->ev_k0) cudaEventDestroy(c->ev_k0);
    if (c->ev_k1) cudaEventDestroy(c->ev_k1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

int dmf_set_camera(dmf_ctx* c, const float K[9], int height, int width) {
    if (!c) return fail("null context");
    if (height < 1 || width < 1) return fail("bad image size %dx%d", height, width);
    std::memcpy(c->K, K, sizeof c->K); c->H = height; c->W = width; c->cam_set = true;
    return 0;
}

int dmf_upload_volume(dmf_ctx* c, const double bounds[6], const double delta[3], const int dim[3],
                      const uint64_t* ids, size_t n_occ, const uint32_t* noff, const float* normals) {
    if (!c) return fail("null context");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    return upload_volume(c, bounds, delta, dim, ids, n_occ, noff, normals);
}

int dmf_volume_from_points(dmf_ctx* c, const double bounds[6], const int dims[3], const float* xyz, const float* normals, size_t n) {
    if (!c) return fail("null context");
    for (int a = 0; a < 3; a++) if (dims[a] < 1) return fail("bad dims");
    HostVolume hv;
    hv.construct(bounds, dims);
    hv.integrate(xyz, normals, n);
    std::vector<uint32_t> noff(hv.occupied.size() + 1, 0);
    std::vector<float> flat;
    for (size_t i = 0; i < hv.occupied.size(); i++) {
        noff[i] = (uint32_t)(flat.size() / 3);
        flat.insert(flat.end(), hv.normals[i].begin(), hv.normals[i].end());
    }
    noff[hv.occupied.size()] = (uint32_t)(flat.size() / 3);
    return dmf_upload_volume(c, bounds, hv.delta, hv.dim, hv.occupied.data(), hv.occupied.size(), noff.data(), flat.data());
}

int dmf_volume_info(dmf_ctx* c, int dims[3], double deltas[3], double* voxel_size, size_t* n_occ, size_t* n_normals) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    for (int a = 0; a < 3; a++) { if (dims) dims[a] = c->vol.dim[a]; if (deltas) deltas[a] = c->vol.delta[a]; }
    if (voxel_size) *voxel_size = c->voxel_size;
    if (n_occ) *n_occ = c->n_occ;
    if (n_normals) *n_normals = c->n_normals;
    return 0;
}
int dmf_volume_get_occupied(dmf_ctx* c, uint64_t* ids) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    DMF_TRY(ensure_host_mirror(c));
    std::copy(c->h_occ.begin(), c->h_occ.end(), ids);
    return 0;
}
int dmf_volume_get_normals(dmf_ctx* c, uint32_t* offsets, float* normals) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    DMF_TRY(ensure_host_mirror(c));
    std::copy(c->h_noff.begin(), c->h_noff.end(), offsets);
    std::copy(c->h_normals.begin(), c->h_normals.end(), normals);
    return 0;
}

int dmf_clear_marks(dmf_ctx* c) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaMemsetAsync(c->d_view_mark.p, 0, std::max<size_t>(c->n_occ, 1) * 4, c->stream));
    DMF_CUDA(cudaMemsetAsync(c->d_good_bits.p, 0, ((c->n_occ + 63) / 64 + 1) * 8, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int dmf_download_marks(dmf_ctx* c, int32_t* view, uint8_t* good, size_t n) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    if (n != c->n_occ) return fail("dmf_download_marks: caller has %zu voxels, the uploaded volume %zu", n, c->n_occ);
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    if (view && c->n_occ) DMF_CUDA(cudaMemcpy(view, c->d_view_mark.p, c->n_occ * 4, cudaMemcpyDeviceToHost));
    if (good && c->n_occ) {
        std::vector<uint32_t> bits((c->n_occ + 31) / 32);
        DMF_CUDA(cudaMemcpy(bits.data(), c->d_good_bits.p, bits.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < c->n_occ; i++) good[i] = (bits[i >> 5] >> (i & 31)) & 1u;
    }
    return 0;
}

int dmf_upload_marks(dmf_ctx* c, const int32_t* view, const uint8_t* good, size_t n) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    if (n != c->n_occ) return fail("dmf_upload_marks: caller has %zu voxels, the uploaded volume %zu", n, c->n_occ);
    if (!c->n_occ) return 0;                       // an empty volume has no marks (std::vector<>(0).data() may be null)
    if (!view || !good) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> bits((c->n_occ + 63) / 64 * 2 + 2, 0);
    for (size_t i = 0; i < c->n_occ; i++) if (good[i]) bits[i >> 5] |= 1u << (i & 31);
    DMF_CUDA(cudaMemcpy(c->d_view_mark.p, view, c->n_occ * 4, cudaMemcpyHostToDevice));
    DMF_CUDA(cudaMemcpy(c->d_good_bits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

// ---- carve mode: the observed-voxel bit grid ------------------------------------------------------------------
size_t dmf_observed_words(dmf_ctx* c) { return c && c->vol_set ? c->n_grid_words : 0; }

int dmf_clear_observed(dmf_ctx* c) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    c->observed_ready = false;
    DMF_TRY(ensure_observed(c, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int dmf_download_observed(dmf_ctx* c, uint32_t* words) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    if (!words) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_TRY(ensure_observed(c, c->stream));
    DMF_CUDA(cudaDeviceSynchronize());
    DMF_CUDA(cudaMemcpy(words, c->d_observed.p, c->n_grid_words * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dmf_observed_dev(dmf_ctx* c, void** d_words) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    if (!d_words) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_TRY(ensure_observed(c, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    *d_words = c->d_observed.p;
    return 0;
}

int dmf_observed_counts(dmf_ctx* c, uint64_t out[3]) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    if (!out) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_TRY(ensure_observed(c, c->stream));
    DMF_CUDA(cudaDeviceSynchronize());
    DMF_TRY(c->d_misc[3].reserve(16));
    DMF_CUDA(cudaMemsetAsync(c->d_misc[3].p, 0, 16, c->stream));
    k_observed_counts<<<blocks_for(c->n_grid_words, 256), 256, 0, c->stream>>>(c->d_observed.as<unsigned>(), c->d_bricks.as<unsigned>(), c->n_grid_words, c->d_misc[3].as<u64>());
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    uint64_t h[2];
    DMF_CUDA(cudaMemcpyAsync(h, c->d_misc[3].p, 16, cudaMemcpyDeviceToHost, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[0] - h[1];
    return 0;
}

size_t dmf_visibility_words(dmf_ctx* c) { return c && c->vol_set ? (c->n_occ + 63) / 64 : 0; }

int dmf_forward_dev(dmf_ctx* c, const dmf_forward_params* p, const float* d_poses, int n_views, const dmf_forward_out* d_out, void* stream) {
    if (!c || !p || !d_out) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    if (d_out->ids || d_out->ids_offsets) return fail("dmf_forward_dev does not produce id lists; use dmf_forward or the visibility bitset");
    FwdPlan pl; DMF_TRY(plan_forward(c, p, pl));
    cudaStream_t st = pick_stream(c, stream);
    DMF_TRY(order_after_previous(c, st));
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    DMF_TRY(enqueue_forward(c, p, pl, d_poses, n_views, p->view_id0, *d_out, nullptr, nullptr, nullptr, st));
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    return mark_last(c, st);
}

int dmf_forward(dmf_ctx* c, const dmf_forward_params* p, const float* poses, int n_views, const dmf_forward_out* out) {
    if (!c || !p || !out || (!poses && n_views > 0)) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    FwdPlan pl; DMF_TRY(plan_forward(c, p, pl));
    const bool want_ids = out->ids_offsets != nullptr;
    if (want_ids && p->mode != DMF_MODE_POINTS && p->mode != DMF_MODE_GOOD_POINTS) return fail("id lists exist only for POINTS / GOOD_POINTS");
    const size_t HW = (size_t)c->H * c->W;
    const size_t vw = (c->n_occ + 63) / 64;
    const int Wc = (c->W + pl.cstride - 1) / pl.cstride, Hc = (c->H + pl.rstride - 1) / pl.rstride;
    const size_t R = (size_t)Wc * Hc;
    const int S = (1000 - pl.z0 + p->zdelta - 1) / p->zdelta;
    if (want_ids && (R > (1u << 21) || S > 1024)) return fail("id lists need <= 2^21 cast pixels and <= 1024 z-planes (got %zu, %d)", R, S);
    if (want_ids) out->ids_offsets[0] = 0;
    // views per chunk: keep the per-chunk device footprint around 1 GiB
    size_t per_view = 48 + (out->depth_u16 ? HW * 2 : 0) + (out->depth_mm ? HW * 4 : 0) + (out->points ? HW * 12 : 0) + (out->hit_voxel ? HW * 8 : 0) +
                      (out->visibility ? vw * 8 : 0) + 8 + (want_ids ? c->n_occ * 4 + R * 20 : 0);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)n_views, (size_t)4096, ((size_t)1 << 30) / per_view}));
    if (!want_ids && n_views >= 16) chunk = std::min(chunk, (n_views + 7) / 8);      // >= 8 chunks: D2H of one overlaps the march of the next
    else if (!want_ids && n_views >= 8) chunk = std::min(chunk, (n_views + 3) / 4);
    if (const char* e = std::getenv("DMF_FWD_CHUNKS")) { const int nc = std::atoi(e); if (nc >= 1 && !want_ids) chunk = std::max(1, (n_views + nc - 1) / nc); }   // tuning aid
    cudaStream_t st = c->stream, cs = c->copy_stream;
    DMF_TRY(order_after_previous(c, st));
    if (!want_ids) {
        // Whole batch resident on the device (up to 4 GiB of outputs per pass; HBM has room), marched as a few launches over
        // consecutive view ranges.  Nothing but those launches sits on the compute stream between them, so the GPU never idles;
        // the per-pixel outputs of a finished range travel to the host on the copy stream while the next range is marched.
        size_t pass = std::min<size_t>({(size_t)n_views, (size_t)65535, std::max<size_t>(1, ((size_t)4 << 30) / per_view), (size_t)(4294967295ull / HW)});
        if (vw) pass = std::min<size_t>(pass, (size_t)(4294967295ull / (2 * vw)));
        DMF_CUDA(cudaEventRecord(c->ev_k0, st));
        for (int s0 = 0; s0 < n_views; s0 += (int)pass) {
            const int ns = std::min((int)pass, n_views - s0);
            if (s0 > 0) DMF_CUDA(cudaStreamWaitEvent(st, c->ev_copied[0], 0));       // the device buffers are reused by the next pass
            DevBuf* ob = c->d_out[0];
            dmf_forward_out d{};
            DMF_TRY(c->d_poses[0].reserve((size_t)ns * 48));
            DMF_CUDA(cudaMemcpyAsync(c->d_poses[0].p, poses + 12 * (size_t)s0, (size_t)ns * 48, cudaMemcpyHostToDevice, st));
            if (out->depth_mm) { DMF_TRY(ob[0].reserve(ns * HW * 4)); d.depth_mm = ob[0].as<int32_t>(); }
            if (out->depth_u16) { DMF_TRY(ob[7].reserve(ns * HW * 2)); d.depth_u16 = ob[7].as<uint16_t>(); }
            if (out->points) { DMF_TRY(ob[1].reserve(ns * HW * 12)); d.points = ob[1].as<float>(); }
            if (out->hit_voxel) { DMF_TRY(ob[2].reserve(ns * HW * 8)); d.hit_voxel = ob[2].as<uint64_t>(); }
            if (out->visibility && vw) { DMF_TRY(ob[3].reserve(ns * vw * 8)); d.visibility = ob[3].as<uint64_t>(); }
            DMF_TRY(ob[4].reserve((size_t)ns * 4)); d.found_any = ob[4].as<int32_t>();
            if (p->mode == DMF_MODE_MINIMUM) { DMF_TRY(ob[5].reserve((size_t)ns * 4)); d.min_depth = ob[5].as<int32_t>(); }
            int n_sub = ns >= 16 ? 8 : (ns >= 8 ? 4 : 1);      // measured on B200 (128 VGA views): 8 ranges beat 4 and 16
            if (const char* e = std::getenv("DMF_FWD_CHUNKS")) { const int nc = std::atoi(e); if (nc >= 1) n_sub = nc; }      // tuning aid
            const bool per_pixel = out->depth_mm || out->depth_u16 || out->points || out->hit_voxel;
            if (!per_pixel) n_sub = 1;
            const std::function<int(int, int, cudaStream_t)> after_sub = [&](int v0, int nv, cudaStream_t launched_on) -> int {
                if (!per_pixel) return 0;
                DMF_CUDA(cudaEventRecord(c->ev_compute[0], launched_on));
                DMF_CUDA(cudaStreamWaitEvent(cs, c->ev_compute[0], 0));
                const size_t g0 = (size_t)(s0 + v0);
                if (out->depth_mm) DMF_CUDA(cudaMemcpyAsync(out->depth_mm + g0 * HW, d.depth_mm + v0 * HW, nv * HW * 4, cudaMemcpyDeviceToHost, cs));
                if (out->depth_u16) DMF_CUDA(cudaMemcpyAsync(out->depth_u16 + g0 * HW, d.depth_u16 + v0 * HW, nv * HW * 2, cudaMemcpyDeviceToHost, cs));
                if (out->points) DMF_CUDA(cudaMemcpyAsync(out->points + g0 * HW * 3, d.points + v0 * HW * 3, nv * HW * 12, cudaMemcpyDeviceToHost, cs));
                if (out->hit_voxel) DMF_CUDA(cudaMemcpyAsync(out->hit_voxel + g0 * HW, d.hit_voxel + v0 * HW, nv * HW * 8, cudaMemcpyDeviceToHost, cs));
                return 0;
            };
            DMF_TRY(enqueue_forward(c, p, pl, c->d_poses[0].as<float>(), ns, p->view_id0 + s0, d, nullptr, nullptr, nullptr, st, (ns + n_sub - 1) / n_sub, &after_sub));
            // per-view results: complete only after the last launch (and the MINIMUM / CLASSIFY finishing kernels)
            DMF_CUDA(cudaEventRecord(c->ev_compute[1], st));
            DMF_CUDA(cudaStreamWaitEvent(cs, c->ev_compute[1], 0));
            if (out->visibility && vw) DMF_CUDA(cudaMemcpyAsync(out->visibility + (size_t)s0 * vw, d.visibility, ns * vw * 8, cudaMemcpyDeviceToHost, cs));
            if (out->found_any) DMF_CUDA(cudaMemcpyAsync(out->found_any + s0, d.found_any, (size_t)ns * 4, cudaMemcpyDeviceToHost, cs));
            if (out->min_depth && d.min_depth) DMF_CUDA(cudaMemcpyAsync(out->min_depth + s0, d.min_depth, (size_t)ns * 4, cudaMemcpyDeviceToHost, cs));
            DMF_CUDA(cudaEventRecord(c->ev_copied[0], cs));
        }
        DMF_CUDA(cudaEventRecord(c->ev_k1, st));
        c->timed = true;
        DMF_CUDA(cudaStreamSynchronize(st));
        DMF_CUDA(cudaStreamSynchronize(cs));
        return 0;
    }
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    int64_t ids_total = 0;
    int n_chunks = 0;
    bool any_copy = false;
    for (int v0 = 0; v0 < n_views; v0 += chunk, n_chunks++) {
        const int nv = std::min(chunk, n_views - v0);
        const int b = n_chunks & 1;
        if (n_chunks >= 2) DMF_CUDA(cudaStreamWaitEvent(st, c->ev_copied[b], 0));   // buffer set b is free again
        DMF_TRY(c->d_poses[b].reserve((size_t)nv * 48));
        DMF_CUDA(cudaMemcpyAsync(c->d_poses[b].p, poses + 12 * (size_t)v0, (size_t)nv * 48, cudaMemcpyHostToDevice, st));
        dmf_forward_out d{};
        DevBuf* ob = c->d_out[b];
        if (out->depth_mm) { DMF_TRY(ob[0].reserve(nv * HW * 4)); d.depth_mm = ob[0].as<int32_t>(); }
        if (out->depth_u16) { DMF_TRY(ob[7].reserve(nv * HW * 2)); d.depth_u16 = ob[7].as<uint16_t>(); }
        if (out->points) { DMF_TRY(ob[1].reserve(nv * HW * 12)); d.points = ob[1].as<float>(); }
        if (out->hit_voxel) { DMF_TRY(ob[2].reserve(nv * HW * 8)); d.hit_voxel = ob[2].as<uint64_t>(); }
        if (out->visibility && vw) { DMF_TRY(ob[3].reserve(nv * vw * 8)); d.visibility = ob[3].as<uint64_t>(); }
        DMF_TRY(ob[4].reserve((size_t)nv * 4)); d.found_any = ob[4].as<int32_t>();
        if (p->mode == DMF_MODE_MINIMUM) { DMF_TRY(ob[5].reserve((size_t)nv * 4)); d.min_depth = ob[5].as<int32_t>(); }
        unsigned *fk = nullptr, *rk = nullptr; int* ro = nullptr;
        if (want_ids) {
            DMF_TRY(c->d_first_key.reserve(std::max<size_t>((size_t)nv * c->n_occ, 1) * 4)); DMF_TRY(c->d_ray_key.reserve(nv * R * 4)); DMF_TRY(c->d_ray_occ.reserve(nv * R * 4));
            fk = c->d_first_key.as<unsigned>(); rk = c->d_ray_key.as<unsigned>(); ro = c->d_ray_occ.as<int>();
        }
        if (!want_ids) DMF_TRY(enqueue_forward(c, p, pl, c->d_poses[b].as<float>(), nv, p->view_id0 + v0, d, fk, rk, ro, st));
        // (declared here: used by the host code after the synchronisation below)
        long long cap_dev = 0; size_t off_bytes = 0, found_bytes = 0, first_ids = 0; char* hs = nullptr;
        if (want_ids) {
            // The id lists leave the device without a host round trip in the middle: counts -> offsets -> gather all on the device,
            // then ONE copy of (offsets, found flags, the first ids) into pinned staging and one synchronisation.  Only a view list
            // longer than the staging area costs a second copy.  (Round 1: D2H of the counts, sync, H2D of the offsets, gather, D2H,
            // sync -- 0.36 ms per single-view call, most of it waiting.)
            DMF_TRY(c->d_tmp_a.reserve(nv * R * 4)); DMF_TRY(c->d_tmp_b.reserve(nv * R * 4)); DMF_TRY(c->d_out_occ.reserve(nv * R * 4));
            DMF_TRY(c->d_n_ids.reserve((size_t)nv * 4)); DMF_TRY(c->d_offsets.reserve((size_t)(nv + 1) * 8));
            const int nb = (int)((R + WIN_BLOCK - 1) / WIN_BLOCK);
            DMF_TRY(c->d_misc[0].reserve((size_t)nv * nb * 4));
            cap_dev = (long long)nv * (long long)std::min<size_t>(R, std::max<size_t>(c->n_occ, 1));     // a view cannot return more ids than rays or voxels
            DMF_TRY(c->d_ids.reserve((size_t)std::max<long long>(cap_dev, 1) * 8));
            const int nblk = (int)((std::min<size_t>(R, std::max<size_t>(c->n_occ, 1)) + ORD_TILE - 1) / ORD_TILE);    // winners per view <= min(rays, voxels)
            DMF_TRY(c->d_misc[2].reserve((size_t)nv * 32 * nblk * 4));
            // staging layout: [offsets (nv+1) x 8][found_any nv x 4, padded to 8][first ids]
            off_bytes = (size_t)(nv + 1) * 8; found_bytes = ((size_t)nv * 4 + 7) / 8 * 8;
            first_ids = (size_t)std::min<long long>(cap_dev, 128 * 1024);
            DMF_TRY(c->stage.reserve(off_bytes + found_bytes + first_ids * 8));
            hs = (char*)c->stage.p;
            // everything from here to the copies into the staging area only ENQUEUES on st: as one captured graph for single-view calls
            auto enqueue_all = [&]() -> int {
                DMF_TRY(enqueue_forward(c, p, pl, c->d_poses[b].as<float>(), nv, p->view_id0 + v0, d, fk, rk, ro, st));
                k_win_count<<<dim3(nb, nv), WIN_THREADS, 0, st>>>(rk, ro, fk, c->d_misc[0].as<unsigned>(), (int)R, (int)c->n_occ, nb);
                // (offsets of the compaction, of the radix passes and of the id lists are folded into their consumers: a single-view call
                // is bound by launches and the gaps between them, not by work)
                k_win_compact<<<dim3(nb, nv), WIN_THREADS, 0, st>>>(rk, ro, fk, nullptr, c->d_tmp_a.as<unsigned>(), (int)R, (int)c->n_occ, nb, c->d_misc[0].as<unsigned>(), c->d_n_ids.as<int>());
                {   // discovery order: stable 2 x 5-bit radix sort of the compacted keys on their z-plane (bits 21..30), multi-block
                    unsigned* hist = c->d_misc[2].as<unsigned>();
                    const bool fold = nblk <= 512;
                    k_ord_hist<<<dim3(nblk, nv), ORD_TILE, 0, st>>>(c->d_tmp_a.as<unsigned>(), c->d_n_ids.as<int>(), hist, (int)R, nblk, 21);
                    if (fold) k_ord_scatter<false, true><<<dim3(nblk, nv), ORD_TILE, 0, st>>>(c->d_tmp_a.as<unsigned>(), c->d_n_ids.as<int>(), hist, c->d_tmp_b.as<unsigned>(), nullptr, nullptr, (int)R, nblk, 21);
                    else { k_ord_scan<<<nv, 1024, 0, st>>>(hist, nblk); k_ord_scatter<false, false><<<dim3(nblk, nv), ORD_TILE, 0, st>>>(c->d_tmp_a.as<unsigned>(), c->d_n_ids.as<int>(), hist, c->d_tmp_b.as<unsigned>(), nullptr, nullptr, (int)R, nblk, 21); }
                    k_ord_hist<<<dim3(nblk, nv), ORD_TILE, 0, st>>>(c->d_tmp_b.as<unsigned>(), c->d_n_ids.as<int>(), hist, (int)R, nblk, 26);
                    if (fold) k_ord_scatter<true, true><<<dim3(nblk, nv), ORD_TILE, 0, st>>>(c->d_tmp_b.as<unsigned>(), c->d_n_ids.as<int>(), hist, nullptr, ro, c->d_out_occ.as<int>(), (int)R, nblk, 26);
                    else { k_ord_scan<<<nv, 1024, 0, st>>>(hist, nblk); k_ord_scatter<true, false><<<dim3(nblk, nv), ORD_TILE, 0, st>>>(c->d_tmp_b.as<unsigned>(), c->d_n_ids.as<int>(), hist, nullptr, ro, c->d_out_occ.as<int>(), (int)R, nblk, 26); }
                    c->launches += fold ? 3 : 5;
                }
                k_gather_ids<<<dim3(32, nv), 256, 0, st>>>(c->d_out_occ.as<int>(), c->d_offsets.as<long long>(), c->vol.occ_ids, c->d_ids.as<u64>(), (int)R, cap_dev, c->d_n_ids.as<int>());
                c->launches += 3;
                DMF_CUDA(cudaGetLastError());
                DMF_CUDA(cudaMemcpyAsync(hs, c->d_offsets.p, off_bytes, cudaMemcpyDeviceToHost, st));
                DMF_CUDA(cudaMemcpyAsync(hs + off_bytes, d.found_any, (size_t)nv * 4, cudaMemcpyDeviceToHost, st));
                if (first_ids) DMF_CUDA(cudaMemcpyAsync(hs + off_bytes + found_bytes, c->d_ids.p, first_ids * 8, cudaMemcpyDeviceToHost, st));
                return 0;
            };
            const bool per_pixel_or_vis = out->depth_mm || out->depth_u16 || out->points || out->hit_voxel || (out->visibility && vw);
            if (n_views == 1 && !per_pixel_or_vis) {
                // the shape of the drop-in's per-view calls: replayed as a captured graph from the second identical call on
                struct { int mode, zdelta, sparse, flags, fmt, H, W, bytes_built; float K[9]; unsigned long long n_occ, epoch, R; const void* ids_out; } k;
                std::memset(&k, 0, sizeof k);
                k.mode = p->mode; k.zdelta = p->zdelta; k.sparse = p->sparse; k.flags = p->flags; k.fmt = pl.grid_format; k.H = c->H; k.W = c->W; k.bytes_built = c->bytes_built ? 1 : 0;
                std::memcpy(k.K, c->K, sizeof k.K); k.n_occ = c->n_occ; k.epoch = c->volume_epoch; k.R = R;
                // whatever may allocate or synchronise happens before a capture can start (another call may have replaced the tables since)
                DMF_TRY(ensure_tables(c, pl.z0, p->zdelta, pl.cstride, pl.rstride, st));
                if (pl.grid_format == DMF_GRID_BYTE) DMF_TRY(ensure_bytes(c, st));
                if (p->flags & DMF_FWD_CARVE) DMF_TRY(ensure_observed(c, st));
                DMF_TRY(run_maybe_graphed(c, c->graph_fwd_ids, fnv1a(&k, sizeof k), st, enqueue_all));
            } else DMF_TRY(enqueue_all());
            DMF_CUDA(cudaStreamSynchronize(st));
            const long long* offs = (const long long*)hs;
            const long long total = offs[nv];
            if (total > cap_dev) return fail("internal: %lld ids exceed the device bound %lld", total, cap_dev);
            if ((size_t)(ids_total + total) > out->ids_capacity || (!out->ids && total > 0)) return fail("ids_capacity %zu too small (need >= %lld)", out->ids_capacity, (long long)(ids_total + total));
            if (total > 0) {
                const size_t head = (size_t)std::min<long long>(total, (long long)first_ids);
                std::memcpy(out->ids + ids_total, hs + off_bytes + found_bytes, head * 8);
                if ((size_t)total > head) DMF_CUDA(cudaMemcpy(out->ids + ids_total + head, c->d_ids.as<u64>() + head, ((size_t)total - head) * 8, cudaMemcpyDeviceToHost));
            }
            for (int i = 0; i < nv; i++) out->ids_offsets[v0 + i + 1] = ids_total + offs[i + 1];
            if (out->found_any) std::memcpy(out->found_any + v0, hs + off_bytes, (size_t)nv * 4);
            ids_total += total;
        }
        DMF_CUDA(cudaEventRecord(c->ev_compute[b], st));
        DMF_CUDA(cudaStreamWaitEvent(cs, c->ev_compute[b], 0));
        if (out->depth_mm) DMF_CUDA(cudaMemcpyAsync(out->depth_mm + v0 * HW, d.depth_mm, nv * HW * 4, cudaMemcpyDeviceToHost, cs));
        if (out->depth_u16) DMF_CUDA(cudaMemcpyAsync(out->depth_u16 + v0 * HW, d.depth_u16, nv * HW * 2, cudaMemcpyDeviceToHost, cs));
        if (out->points) DMF_CUDA(cudaMemcpyAsync(out->points + v0 * HW * 3, d.points, nv * HW * 12, cudaMemcpyDeviceToHost, cs));
        if (out->hit_voxel) DMF_CUDA(cudaMemcpyAsync(out->hit_voxel + v0 * HW, d.hit_voxel, nv * HW * 8, cudaMemcpyDeviceToHost, cs));
        if (out->visibility && vw) DMF_CUDA(cudaMemcpyAsync(out->visibility + v0 * vw, d.visibility, nv * vw * 8, cudaMemcpyDeviceToHost, cs));
        if (out->found_any && !want_ids) DMF_CUDA(cudaMemcpyAsync(out->found_any + v0, d.found_any, (size_t)nv * 4, cudaMemcpyDeviceToHost, cs));
        if (out->min_depth && d.min_depth) DMF_CUDA(cudaMemcpyAsync(out->min_depth + v0, d.min_depth, (size_t)nv * 4, cudaMemcpyDeviceToHost, cs));
        any_copy = any_copy || out->depth_mm || out->depth_u16 || out->points || out->hit_voxel || (out->visibility && vw) || (out->found_any && !want_ids) || (out->min_depth && d.min_depth);
        DMF_CUDA(cudaEventRecord(c->ev_copied[b], cs));
    }
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    if (!want_ids || n_chunks > 1) DMF_CUDA(cudaStreamSynchronize(st));      // (the id path has just synchronised its only chunk)
    if (any_copy || n_chunks > 2) DMF_CUDA(cudaStreamSynchronize(cs));
    return 0;
}

int dmf_counters(dmf_ctx* c, uint64_t out[DMF_CNT_COUNT]) {
    if (!c) return fail("null context");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    std::vector<uint64_t> all((size_t)DMF_COUNTER_SLOTS * DMF_COUNTER_STRIDE);
    DMF_CUDA(cudaMemcpy(all.data(), c->d_counters.p, all.size() * 8, cudaMemcpyDeviceToHost));
    for (int j = 0; j < DMF_CNT_COUNT; j++) { out[j] = 0; for (int s = 0; s < DMF_COUNTER_SLOTS; s++) out[j] += all[(size_t)s * DMF_COUNTER_STRIDE + j]; }
    out[DMF_CNT_LAUNCHES] = c->launches;
#ifdef DMF_LINE_STATS
    // diagnostic build: k_forward_line logs blocks / block-cycles / warp-cycles per SM into the unused tail of each slot
    for (int s = 0; s < 160; s++) {
        const uint64_t* q = &all[(size_t)s * DMF_COUNTER_STRIDE];
        if (q[12]) std::fprintf(stderr, "sm %3d blocks %7llu  warp-cycles/warp %8.0f  prologue %8.0f  march loop %8.0f (line %8.0f  exact %8.0f)\n", s, (unsigned long long)q[12], q[13] / (4.0 * q[12]),
                                q[11] / (4.0 * q[12]), q[10] / (4.0 * q[12]), q[14] / (4.0 * q[12]), q[15] / (4.0 * q[12]));
    }
#endif
    return 0;
}
int dmf_reset_counters(dmf_ctx* c) {
    if (!c) return fail("null context");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    DMF_CUDA(cudaMemset(c->d_counters.p, 0, DMF_COUNTER_SLOTS * DMF_COUNTER_STRIDE * 8));
    c->launches = 0;
    return 0;
}
int dmf_last_kernel_ms(dmf_ctx* c, float* ms) {
    if (!c || !ms) return fail("null argument");
    if (!c->timed) return fail("no timed call yet");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaEventSynchronize(c->ev_k1));
    DMF_CUDA(cudaEventElapsedTime(ms, c->ev_k0, c->ev_k1));
    return 0;
}
int dmf_last_hot_kernel_ms(dmf_ctx* c, float* ms) {
    if (!c || !ms) return fail("null argument");
    if (!c->hot_timed) return fail("no march launched yet");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaEventSynchronize(c->ev_h1));
    DMF_CUDA(cudaEventElapsedTime(ms, c->ev_h0, c->ev_h1));
    return 0;
}
int dmf_volume_prepare_ms(dmf_ctx* c, float* build_ms, float* bytes_ms) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    DMF_CUDA(cudaSetDevice(c->device));
    if (build_ms) { *build_ms = -1.f; if (c->prepare_timed) { DMF_CUDA(cudaEventSynchronize(c->ev_p1)); DMF_CUDA(cudaEventElapsedTime(build_ms, c->ev_p0, c->ev_p1)); } }
    if (bytes_ms) { *bytes_ms = -1.f; if (c->bytes_built && c->bytes_timed) { DMF_CUDA(cudaEventSynchronize(c->ev_b1)); DMF_CUDA(cudaEventElapsedTime(bytes_ms, c->ev_b0, c->ev_b1)); } }
    return 0;
}
int dmf_prepare_grid(dmf_ctx* c, int grid_format) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    DMF_CUDA(cudaSetDevice(c->device));
    if (grid_format == DMF_GRID_BYTE) return ensure_bytes(c, c->stream);
    if (grid_format != DMF_GRID_BIT) return fail("bad grid_format %d", grid_format);
    return 0;
}
int dmf_synchronize(dmf_ctx* c) {
    if (!c) return fail("null context");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"

#include "dmf_abi_rest.cuh"
#include "dmf_comm.cuh"
