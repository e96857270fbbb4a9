// dmf_b200.cu -- libdmf_b200.so: C ABI (include/dmf_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 (see build.py).
// There is no CPU fallback in this file: every compute entry point launches CUDA kernels or fails.
#include "dmf_host.cuh"
#include "dmf_forward.cuh"
#include "dmf_reverse.cuh"
#include "dmf_distance.cuh"
#include "dmf_integrate.cuh"
#include "dmf_setcover.cuh"
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <functional>
#include <new>

using namespace dmf;

namespace {

inline cudaStream_t pick_stream(dmf_ctx* c, void* s) { return s ? (cudaStream_t)s : c->stream; }
inline unsigned blocks_for(size_t n, unsigned threads, unsigned cap = 148 * 16) {
    size_t b = (n + threads - 1) / threads;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(b, cap));
}

int fill_u32(dmf_ctx* c, cudaStream_t st, void* p, size_t n, unsigned val) {
    if (!n) return 0;
    k_fill_u32<<<blocks_for(n, 256), 256, 0, st>>>((unsigned*)p, n, val);
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    return 0;
}

// ---- volume upload ------------------------------------------------------------------------------------------
int upload_volume(dmf_ctx* c, const double bounds[6], const double delta[3], const int dim[3],
                  const uint64_t* ids, size_t n_occ, const uint32_t* noff, const float* normals) {
    DMF_CUDA(cudaSetDevice(c->device));
    for (int a = 0; a < 3; a++) {
        if (dim[a] < 1 || dim[a] > 2048) return fail("volume dim %d out of range [1,2048] (voxel ids shift y by 20 bits as int, Volume.hpp:146)", dim[a]);
        if (!(delta[a] > 0) || !(bounds[2 * a + 1] > bounds[2 * a])) return fail("degenerate volume bounds/delta on axis %d", a);
    }
    VolDev& v = c->vol;
    std::memcpy(c->bounds, bounds, sizeof c->bounds);
    c->voxel_size = delta[0] * delta[1] * delta[2];
    if ((double)(dim[0] + 1) * (dim[1] + 1) * (dim[2] + 1) >= 4294967296.0) return fail("volume %dx%dx%d too large for the 32-bit linear voxel index", dim[0], dim[1], dim[2]);
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
    const size_t nbits = (size_t)v.pdim[0] * v.pdim[1] * v.pdim[2];
    const size_t nwords = ((nbits + 31) / 32 + 7) / 8 * 8;           // whole 8-word (256-bit) rank blocks
    const size_t nmacro = (size_t)v.mdim[0] * v.mdim[1] * v.mdim[2];
    std::vector<uint32_t> words(nwords, 0), prefix(nwords, 0), macro((nmacro + 31) / 32, 0);
    auto locate = [&](uint64_t id, size_t& idx) -> bool {
        const uint64_t mask = (1u << 20) - 1;
        long long x = (long long)(id >> 40), y = (long long)((id >> 20) & mask), z = (long long)(id & mask);   // getVoxelCoords :158-165
        if (x >= dim[0] || y >= dim[1] || z >= dim[2]) return false;
        idx = ((size_t)x * v.pdim[1] + (size_t)y) * v.pdim[2] + (size_t)z;
        size_t m = ((size_t)(x >> 3) * v.mdim[1] + (size_t)(y >> 3)) * v.mdim[2] + (size_t)(z >> 3);
        macro[m >> 5] |= 1u << (m & 31);
        return true;
    };
    for (size_t i = 0; i < n_occ; i++) {
        size_t idx;
        if (!locate(ids[i], idx)) return fail("occupied id %llu (#%zu) outside the %dx%dx%d grid", (unsigned long long)ids[i], i, dim[0], dim[1], dim[2]);
        if ((words[idx >> 5] >> (idx & 31)) & 1u) return fail("duplicate occupied id %llu (#%zu)", (unsigned long long)ids[i], i);
        words[idx >> 5] |= 1u << (idx & 31);
    }
    std::vector<uint32_t>& word_rank = prefix;
    uint32_t run = 0;
    for (size_t w = 0; w < nwords; w++) { prefix[w] = run; run += (uint32_t)__builtin_popcount(words[w]); }
    std::vector<uint32_t> rank2occ(std::max<size_t>(n_occ, 1));
    for (size_t i = 0; i < n_occ; i++) {
        size_t idx; locate(ids[i], idx);
        rank2occ[word_rank[idx >> 5] + __builtin_popcount(words[idx >> 5] & ((1u << (idx & 31)) - 1u))] = (uint32_t)i;
    }
    // Chebyshev distance (in macro cells) from every macro cell to the nearest "blocked" cell: one that holds an occupied
    // voxel, is not entirely inside [0,dim) on every axis, or lies outside the grid.  Box dilation is separable, so each
    // radius step is three 1-D passes.  clearance = 8*(D-1) - 0.25 voxels for D >= 2, else 0 (see k_forward_skip).
    std::vector<float> clearance(nmacro, 0.0f);
    {
        const int mx = v.mdim[0], my = v.mdim[1], mz = v.mdim[2];
        const int ex = mx + 2, ey = my + 2, ez = mz + 2;                      // one blocked border cell on every side
        auto at = [&](int x, int y, int z) { return ((size_t)x * ey + y) * ez + z; };
        std::vector<uint8_t> cur((size_t)ex * ey * ez, 1), tmp(cur.size());
        std::vector<uint8_t> dist(nmacro, 0);
        for (int x = 0; x < mx; x++) for (int y = 0; y < my; y++) for (int z = 0; z < mz; z++) {
            size_t m = ((size_t)x * my + y) * mz + z;
            bool blocked = ((macro[m >> 5] >> (m & 31)) & 1u) || 8 * (x + 1) > dim[0] || 8 * (y + 1) > dim[1] || 8 * (z + 1) > dim[2];
            cur[at(x + 1, y + 1, z + 1)] = blocked ? 1 : 0;
        }
        const int kMaxD = 40;
        for (int r = 1; r <= kMaxD; r++) {
            // dilate by one cell along z, then y, then x
            for (int pass = 0; pass < 3; pass++) {
                const size_t stride = pass == 0 ? 1 : (pass == 1 ? (size_t)ez : (size_t)ey * ez);
                for (size_t i = 0; i < cur.size(); i++) {
                    uint8_t c = cur[i];
                    if (!c) { if (i >= stride && cur[i - stride]) c = 1; else if (i + stride < cur.size() && cur[i + stride]) c = 1; }
                    tmp[i] = c;
                }
                // the flat +-stride neighbours wrap across rows only into border cells, which are blocked anyway
                cur.swap(tmp);
            }
            size_t fresh = 0;
            for (int x = 0; x < mx; x++) for (int y = 0; y < my; y++) for (int z = 0; z < mz; z++) {
                size_t m = ((size_t)x * my + y) * mz + z;
                if (cur[at(x + 1, y + 1, z + 1)] && dist[m] == 0) {
                    bool blocked0 = ((macro[m >> 5] >> (m & 31)) & 1u) || 8 * (x + 1) > dim[0] || 8 * (y + 1) > dim[1] || 8 * (z + 1) > dim[2];
                    if (!blocked0) { dist[m] = (uint8_t)r; fresh++; }
                }
            }
            if (!fresh) break;
        }
        for (size_t m = 0; m < nmacro; m++) {
            size_t xm = m / ((size_t)my * mz), ym = (m / mz) % my, zm = m % mz;
            bool blocked0 = ((macro[m >> 5] >> (m & 31)) & 1u) || 8 * ((int)xm + 1) > dim[0] || 8 * ((int)ym + 1) > dim[1] || 8 * ((int)zm + 1) > dim[2];
            int D = blocked0 ? 0 : (dist[m] ? dist[m] : kMaxD + 1);           // never reached: farther than kMaxD
            clearance[m] = D >= 2 ? 8.0f * (float)(D - 1) - 0.25f : 0.0f;
        }
    }
    c->n_occ = n_occ;
    c->h_occ.assign(ids, ids + n_occ);
    c->h_noff.assign(n_occ + 1, 0);
    if (noff) c->h_noff.assign(noff, noff + n_occ + 1);
    c->n_normals = c->h_noff[n_occ];
    c->h_normals.assign(3 * c->n_normals, 0.f);
    if (normals && c->n_normals) c->h_normals.assign(normals, normals + 3 * c->n_normals);

    DMF_TRY(c->d_bricks.reserve(nwords * 4)); DMF_TRY(c->d_prefix.reserve(prefix.size() * 4)); DMF_TRY(c->d_macro.reserve(macro.size() * 4)); DMF_TRY(c->d_clearance.reserve(clearance.size() * 4));
    DMF_TRY(c->d_rank2occ.reserve(rank2occ.size() * 4)); DMF_TRY(c->d_occ_ids.reserve(std::max<size_t>(n_occ, 1) * 8));
    DMF_TRY(c->d_noff.reserve((n_occ + 1) * 4)); DMF_TRY(c->d_normals.reserve(std::max<size_t>(c->h_normals.size(), 1) * 4));
    DMF_TRY(c->d_view_mark.reserve(std::max<size_t>(n_occ, 1) * 4)); DMF_TRY(c->d_first_view.reserve(std::max<size_t>(n_occ, 1) * 4));
    DMF_TRY(c->d_good_bits.reserve(((n_occ + 63) / 64 + 1) * 8));
    cudaStream_t st = c->stream;
    DMF_CUDA(cudaMemcpyAsync(c->d_bricks.p, words.data(), nwords * 4, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_prefix.p, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_macro.p, macro.data(), macro.size() * 4, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_clearance.p, clearance.data(), clearance.size() * 4, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_rank2occ.p, rank2occ.data(), rank2occ.size() * 4, cudaMemcpyHostToDevice, st));
    if (n_occ) DMF_CUDA(cudaMemcpyAsync(c->d_occ_ids.p, c->h_occ.data(), n_occ * 8, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_noff.p, c->h_noff.data(), (n_occ + 1) * 4, cudaMemcpyHostToDevice, st));
    if (c->h_normals.size()) DMF_CUDA(cudaMemcpyAsync(c->d_normals.p, c->h_normals.data(), c->h_normals.size() * 4, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemsetAsync(c->d_view_mark.p, 0, std::max<size_t>(n_occ, 1) * 4, st));
    DMF_CUDA(cudaMemsetAsync(c->d_good_bits.p, 0, ((n_occ + 63) / 64 + 1) * 8, st));
    DMF_TRY(fill_u32(c, st, c->d_first_view.p, std::max<size_t>(n_occ, 1), 0x7fffffffu));
    DMF_CUDA(cudaStreamSynchronize(st));   // host staging vectors die here
    v.bits = c->d_bricks.as<unsigned>(); v.prefix = c->d_prefix.as<unsigned>(); v.rank2occ = c->d_rank2occ.as<unsigned>(); v.macro = c->d_macro.as<unsigned>();
    v.noff = c->d_noff.as<unsigned>(); v.normals = c->d_normals.as<float>(); v.occ_ids = c->d_occ_ids.as<u64>();
    v.bytes = nullptr; v.n_occ = (int)n_occ;
    c->bytes_built = false;
    c->n_grid_words = nwords; c->observed_ready = false;     // a new volume starts unobserved
    c->vol_set = true;
    // float-accumulated axes of the whole-grid loops (RayTracingEngine.hpp:54-56, :509-511)
    for (int a = 0; a < 3; a++) {
        std::vector<float> ax;
        double hi = bounds[2 * a + 1];
        for (float x = (float)bounds[2 * a]; x < hi; x = (float)(x + delta[a])) { ax.push_back(x); if (ax.size() > (1u << 22)) break; }
        c->n_axis[a] = (int)ax.size();
        DMF_TRY(c->d_axis[a].reserve(std::max<size_t>(ax.size(), 1) * 4));
        if (!ax.empty()) DMF_CUDA(cudaMemcpy(c->d_axis[a].p, ax.data(), ax.size() * 4, cudaMemcpyHostToDevice));
    }
    // centroid hashes of the occupied voxels (RayTracingEngine.hpp:151-163), view independent
    DMF_TRY(c->d_centroid_hash.reserve(std::max<size_t>(n_occ, 1) * 8));
    if (n_occ) {
        k_centroid_hash<<<blocks_for(n_occ, 256), 256, 0, st>>>(c->vol, c->d_centroid_hash.as<u64>());
        c->launches++;
        DMF_CUDA(cudaGetLastError());
        DMF_CUDA(cudaStreamSynchronize(st));
    }
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
    DevBuf tmp;
    DMF_TRY(tmp.reserve(n));
    const unsigned nlines = (unsigned)v.pdim[0] * (unsigned)v.pdim[1];
    k_dt_z<<<(nlines + 127) / 128, 128, 0, st>>>(v, c->d_bytes.as<unsigned char>());
    k_dt_axis<1, false><<<blocks_for(n, 256, 148 * 32), 256, 0, st>>>(v, c->d_bytes.as<unsigned char>(), tmp.as<unsigned char>());
    k_dt_axis<0, true><<<blocks_for(n, 256, 148 * 32), 256, 0, st>>>(v, tmp.as<unsigned char>(), c->d_bytes.as<unsigned char>());
    c->launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    tmp.release();
    if (e != cudaSuccess) return fail("distance transform failed: %s", cudaGetErrorString(e));
    v.bytes = c->d_bytes.as<unsigned char>();
    c->bytes_built = true;
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

struct FwdPlan { int z0, cstride, rstride; };
int plan_forward(dmf_ctx* c, const dmf_forward_params* p, FwdPlan& pl) {
    if (!c->cam_set) return fail("dmf_set_camera has not been called");
    if (!c->vol_set) return fail("no volume uploaded");
    if (p->mode < 0 || p->mode > 4) return fail("bad mode %d", p->mode);
    if (p->zdelta < 1) return fail("zdelta must be >= 1 (the reference loops forever on zdelta <= 0)");
    if (p->grid_format != DMF_GRID_BIT && p->grid_format != DMF_GRID_BYTE) return fail("bad grid_format %d", p->grid_format);
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
                    int sub_views = 0, const std::function<int(int, int, cudaStream_t)>* after_sub = nullptr) {
    if (n_views <= 0) return 0;
    if (n_views > 65535) return fail("at most 65535 views per launch (got %d)", n_views);
    DMF_TRY(ensure_tables(c, pl.z0, p->zdelta, pl.cstride, pl.rstride, st));
    if (p->grid_format == DMF_GRID_BYTE) DMF_TRY(ensure_bytes(c, st));
    const bool carve = (p->flags & DMF_FWD_CARVE) != 0;
    if (carve) DMF_TRY(ensure_observed(c, st));
    const size_t HW = (size_t)c->H * c->W;
    const size_t vis_words64 = (c->n_occ + 63) / 64;
    const bool sparse_lattice = pl.cstride > 1;
    if (out.depth_mm && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.depth_mm, 0xFF, n_views * HW * 4, st));
    if (out.depth_u16 && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.depth_u16, 0xFF, n_views * HW * 2, st));
    if (out.hit_voxel && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.hit_voxel, 0xFF, n_views * HW * 8, st));
    if (out.points && sparse_lattice) DMF_CUDA(cudaMemsetAsync(out.points, 0, n_views * HW * 12, st));
    if (out.visibility && vis_words64) DMF_CUDA(cudaMemsetAsync(out.visibility, 0, (size_t)n_views * vis_words64 * 8, st));
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
    a.found_any = out.found_any; a.min_depth = out.min_depth;
    a.first_key = first_key; a.ray_key = ray_key; a.ray_occ = ray_occ;
    a.first_view = c->d_first_view.as<int>(); a.good_bits = c->d_good_bits.as<unsigned>(); a.view_mark = c->d_view_mark.as<int>();
    a.counters = c->d_counters.as<u64>();
    a.observed = carve ? c->d_observed.as<unsigned>() : nullptr;
    a.dcx = c->d_dcx.as<float>(); a.dcy = c->d_dcy.as<float>(); a.clearance = c->d_clearance.as<float>();
    a.dcx_max = c->dcx_max; a.dcy_max = c->dcy_max;
    DMF_TRY(c->d_kstart.reserve((size_t)n_views * 8));
    a.kstart = c->d_kstart.as<int>(); a.veps = c->d_kstart.as<float>() + n_views;
    const bool skip = !(p->flags & DMF_FWD_NO_SKIP);
    const bool byte_skip = skip && p->grid_format == DMF_GRID_BYTE;
    const bool two_probe = (p->flags & DMF_FWD_TWO_PROBE) != 0;
    if (sub_views <= 0 || sub_views > n_views) sub_views = n_views;
    DMF_CUDA(cudaEventRecord(c->ev_h0, st));
    if (byte_skip) { a.view0 = 0; k_view_start<<<(n_views + 127) / 128, 128, 0, st>>>(a, n_views, const_cast<int*>(a.kstart)); c->launches++; }
    const bool split = sub_views < n_views;
    if (split) { DMF_CUDA(cudaEventRecord(c->ev_fork, st)); DMF_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0)); }
    int n_launch = 0;
    for (int v0 = 0; v0 < n_views; v0 += sub_views, n_launch++) {
        const int nv = std::min(sub_views, n_views - v0);
        cudaStream_t ls = (n_launch & 1) ? c->aux_stream : st;
        a.view0 = v0;
        dim3 grid((c->Wc + FWD_TILE_W - 1) / FWD_TILE_W, (c->Hc + FWD_TILE_H - 1) / FWD_TILE_H, nv);
        switch (p->mode) {
            case 0: launch_forward_fmt<0>(a, p->grid_format, skip, two_probe, grid, ls); break;
            case 1: launch_forward_fmt<1>(a, p->grid_format, skip, two_probe, grid, ls); break;
            case 2: launch_forward_fmt<2>(a, p->grid_format, skip, two_probe, grid, ls); break;
            case 3: launch_forward_fmt<3>(a, p->grid_format, skip, two_probe, grid, ls); break;
            default: launch_forward_fmt<4>(a, p->grid_format, skip, two_probe, grid, ls); break;
        }
        c->launches++;
        if (after_sub) DMF_TRY((*after_sub)(v0, nv, ls));
    }
    if (split) { DMF_CUDA(cudaEventRecord(c->ev_join, c->aux_stream)); DMF_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0)); }
    DMF_CUDA(cudaEventRecord(c->ev_h1, st));
    c->hot_timed = true;
    DMF_CUDA(cudaGetLastError());
    if (p->mode == DMF_MODE_CLASSIFY && c->n_occ) {
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

int dmf_version(void) { return 100; }

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
    for (int i = 0; i < 2; i++) {
        DMF_CUDA(cudaEventCreateWithFlags(&c->ev_compute[i], cudaEventDisableTiming));
        DMF_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
    DMF_TRY(c->d_counters.reserve(DMF_COUNTER_SLOTS * DMF_COUNTER_STRIDE * 8));
    DMF_CUDA(cudaMemset(c->d_counters.p, 0, DMF_COUNTER_SLOTS * DMF_COUNTER_STRIDE * 8));
    DMF_CUDA(cudaFuncSetAttribute(k_order_ids, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * ORD_THREADS * 4));
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
                      &c->d_out_occ, &c->d_n_ids, &c->d_offsets, &c->d_ids, &c->d_misc[0], &c->d_misc[1], &c->d_misc[2], &c->d_misc[3], &c->d_counters};
    for (auto* b : bufs) b->release();
    for (int i = 0; i < 2; i++) for (int j = 0; j < 8; j++) c->d_out[i][j].release();
    for (int i = 0; i < 2; i++) { if (c->ev_compute[i]) cudaEventDestroy(c->ev_compute[i]); if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); }
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
    std::copy(c->h_occ.begin(), c->h_occ.end(), ids);
    return 0;
}
int dmf_volume_get_normals(dmf_ctx* c, uint32_t* offsets, float* normals) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
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

int dmf_download_marks(dmf_ctx* c, int32_t* view, uint8_t* good) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
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

int dmf_upload_marks(dmf_ctx* c, const int32_t* view, const uint8_t* good) {
    if (!c || !c->vol_set) return fail("no volume uploaded");
    if (!view || !good) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    if (!c->n_occ) return 0;
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
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    DMF_TRY(enqueue_forward(c, p, pl, d_poses, n_views, p->view_id0, *d_out, nullptr, nullptr, nullptr, st));
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    return 0;
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
        DMF_TRY(enqueue_forward(c, p, pl, c->d_poses[b].as<float>(), nv, p->view_id0 + v0, d, fk, rk, ro, st));
        if (want_ids) {
            DMF_TRY(c->d_tmp_a.reserve(nv * R * 4)); DMF_TRY(c->d_tmp_b.reserve(nv * R * 4)); DMF_TRY(c->d_out_occ.reserve(nv * R * 4));
            DMF_TRY(c->d_n_ids.reserve((size_t)nv * 4)); DMF_TRY(c->d_offsets.reserve((size_t)(nv + 1) * 8));
            const int nb = (int)((R + WIN_BLOCK - 1) / WIN_BLOCK);
            DMF_TRY(c->d_misc[0].reserve((size_t)nv * nb * 4)); DMF_TRY(c->d_misc[1].reserve((size_t)nv * nb * 4));
            k_win_count<<<dim3(nb, nv), WIN_THREADS, 0, st>>>(rk, ro, fk, c->d_misc[0].as<unsigned>(), (int)R, (int)c->n_occ, nb);
            k_win_offsets<<<nv, 1024, 0, st>>>(c->d_misc[0].as<unsigned>(), c->d_misc[1].as<unsigned>(), c->d_n_ids.as<int>(), nb);
            k_win_compact<<<dim3(nb, nv), WIN_THREADS, 0, st>>>(rk, ro, fk, c->d_misc[1].as<unsigned>(), c->d_tmp_a.as<unsigned>(), (int)R, (int)c->n_occ, nb);
            k_order_ids<<<nv, ORD_THREADS, 32 * ORD_THREADS * 4, st>>>(ro, c->d_tmp_a.as<unsigned>(), c->d_tmp_b.as<unsigned>(), c->d_out_occ.as<int>(), c->d_n_ids.as<int>(), (int)R);
            c->launches += 4;
            DMF_CUDA(cudaGetLastError());
            std::vector<int> n_ids(nv);
            DMF_CUDA(cudaMemcpyAsync(n_ids.data(), c->d_n_ids.p, (size_t)nv * 4, cudaMemcpyDeviceToHost, st));
            DMF_CUDA(cudaStreamSynchronize(st));
            std::vector<long long> offs(nv + 1, 0);
            for (int i = 0; i < nv; i++) offs[i + 1] = offs[i] + n_ids[i];
            if ((size_t)(ids_total + offs[nv]) > out->ids_capacity || (!out->ids && offs[nv] > 0)) return fail("ids_capacity %zu too small (need >= %lld)", out->ids_capacity, (long long)(ids_total + offs[nv]));
            if (offs[nv] > 0) {
                DMF_TRY(c->d_ids.reserve((size_t)offs[nv] * 8));
                DMF_CUDA(cudaMemcpyAsync(c->d_offsets.p, offs.data(), (size_t)(nv + 1) * 8, cudaMemcpyHostToDevice, st));
                k_gather_ids<<<dim3(32, nv), 256, 0, st>>>(c->d_out_occ.as<int>(), c->d_offsets.as<long long>(), c->vol.occ_ids, c->d_ids.as<u64>(), (int)R);
                c->launches++;
                DMF_CUDA(cudaGetLastError());
                DMF_CUDA(cudaMemcpyAsync(out->ids + ids_total, c->d_ids.p, (size_t)offs[nv] * 8, cudaMemcpyDeviceToHost, st));
                DMF_CUDA(cudaStreamSynchronize(st));
            }
            for (int i = 0; i < nv; i++) out->ids_offsets[v0 + i + 1] = ids_total + offs[i + 1];
            ids_total += offs[nv];
        }
        DMF_CUDA(cudaEventRecord(c->ev_compute[b], st));
        DMF_CUDA(cudaStreamWaitEvent(cs, c->ev_compute[b], 0));
        if (out->depth_mm) DMF_CUDA(cudaMemcpyAsync(out->depth_mm + v0 * HW, d.depth_mm, nv * HW * 4, cudaMemcpyDeviceToHost, cs));
        if (out->depth_u16) DMF_CUDA(cudaMemcpyAsync(out->depth_u16 + v0 * HW, d.depth_u16, nv * HW * 2, cudaMemcpyDeviceToHost, cs));
        if (out->points) DMF_CUDA(cudaMemcpyAsync(out->points + v0 * HW * 3, d.points, nv * HW * 12, cudaMemcpyDeviceToHost, cs));
        if (out->hit_voxel) DMF_CUDA(cudaMemcpyAsync(out->hit_voxel + v0 * HW, d.hit_voxel, nv * HW * 8, cudaMemcpyDeviceToHost, cs));
        if (out->visibility && vw) DMF_CUDA(cudaMemcpyAsync(out->visibility + v0 * vw, d.visibility, nv * vw * 8, cudaMemcpyDeviceToHost, cs));
        if (out->found_any) DMF_CUDA(cudaMemcpyAsync(out->found_any + v0, d.found_any, (size_t)nv * 4, cudaMemcpyDeviceToHost, cs));
        if (out->min_depth && d.min_depth) DMF_CUDA(cudaMemcpyAsync(out->min_depth + v0, d.min_depth, (size_t)nv * 4, cudaMemcpyDeviceToHost, cs));
        DMF_CUDA(cudaEventRecord(c->ev_copied[b], cs));
    }
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    DMF_CUDA(cudaStreamSynchronize(st));
    DMF_CUDA(cudaStreamSynchronize(cs));
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
        if (q[12]) std::fprintf(stderr, "sm %3d blocks %7llu  warp-cycles/warp %8.0f  line %8.0f  exact %8.0f\n", s, (unsigned long long)q[12], q[13] / (4.0 * q[12]), q[14] / (4.0 * q[12]), q[15] / (4.0 * q[12]));
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
int dmf_synchronize(dmf_ctx* c) {
    if (!c) return fail("null context");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"

#include "dmf_abi_rest.cuh"
