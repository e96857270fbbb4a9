// dmf_abi_rest.cuh -- C ABI entry points for the reverse march, the z-buffer, set cover and the OR combine.
// Included at the end of dmf_b200.cu (same translation unit, same helpers).
#pragma once

namespace {

struct RevScratch { unsigned* vis; unsigned* unocc; int* found; };

int fill_rev_args(dmf_ctx* c, RevArgs& a) {
    if (!c->cam_set) return fail("dmf_set_camera has not been called");
    if (!c->vol_set) return fail("no volume uploaded");
    a.vol = c->vol; a.angle = c->angle;
    a.fx = (double)c->K[0]; a.cx = (double)c->K[2]; a.fy = (double)c->K[4]; a.cy = (double)c->K[5];
    a.H = c->H; a.W = c->W;
    a.centroid_hash = c->d_centroid_hash.as<u64>();
    for (int i = 0; i < 3; i++) { a.ax[i] = c->d_axis[i].as<float>(); a.nax[i] = c->n_axis[i]; }
    a.vis = nullptr; a.unocc = nullptr; a.vis_words32 = (int)(((c->n_occ + 63) / 64) * 2);
    a.vis_stride32 = (unsigned)a.vis_words32; std::memset(&a.pub, 0, sizeof a.pub);
    a.found_any = nullptr; a.viz = 0;
    a.view_mark = c->d_view_mark.as<int>(); a.good_bits = c->d_good_bits.as<unsigned>();
    a.emit_list = nullptr; a.emit_count = nullptr; a.emit_cap = 0; a.zbuf = nullptr;
    a.counters = c->d_counters.as<u64>();
    a.step_cap = 1000000;
    a.pnyz = (unsigned)c->vol.pdim[1] * (unsigned)c->vol.pdim[2];
    a.bias = 0x4B400000u * (a.pnyz + (unsigned)c->vol.pdim[2] + 1u);
    return 0;
}

// enqueue the reverse march of n_views poses (device) into device outputs
int enqueue_reverse(dmf_ctx* c, int fast, int viz, const float* d_poses, int n_views, unsigned* d_vis, unsigned* d_unocc, int* d_found,
                    u64* d_emit_list, unsigned* d_emit_count, unsigned emit_cap, cudaStream_t st,
                    unsigned vis_stride32 = 0, const PubTable* pub = nullptr) {
    if (n_views <= 0) return 0;
    if (c->reverse_format == DMF_GRID_BYTE && c->vol_set) DMF_TRY(ensure_bytes(c, st));
    if (n_views > 65535) return fail("at most 65535 views per launch (got %d)", n_views);
    RevArgs a; DMF_TRY(fill_rev_args(c, a));
    DMF_TRY(c->d_inv_poses.reserve((size_t)n_views * 48));
    k_invert_poses<<<(n_views + 127) / 128, 128, 0, st>>>(d_poses, c->d_inv_poses.as<float>(), n_views);
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    const size_t vw = (c->n_occ + 63) / 64;
    if (d_vis && vw && !vis_stride32) DMF_CUDA(cudaMemsetAsync(d_vis, 0, (size_t)n_views * vw * 8, st));   // (a sharded sweep zeroes its interleaved rows itself)
    if (d_unocc && vw) DMF_CUDA(cudaMemsetAsync(d_unocc, 0, (size_t)n_views * vw * 8, st));
    if (d_found) DMF_CUDA(cudaMemsetAsync(d_found, 0, (size_t)n_views * 4, st));
    if (d_emit_count) DMF_CUDA(cudaMemsetAsync(d_emit_count, 0, (size_t)n_views * 4, st));
    a.poses = d_poses; a.inv_poses = c->d_inv_poses.as<float>();
    a.vis = d_vis; a.unocc = d_unocc; a.found_any = d_found; a.viz = viz;
    if (vis_stride32) a.vis_stride32 = vis_stride32;
    if (pub) a.pub = *pub;
    a.emit_list = d_emit_list; a.emit_count = d_emit_count; a.emit_cap = emit_cap;
    if (!c->capturing) DMF_CUDA(cudaEventRecord(c->ev_h0, st));
    if (fast) {
        if (!c->n_occ) return 0;
        dim3 grid((unsigned)((c->n_occ + 127) / 128), n_views);
        // DMF_REVERSE_POOL=1 selects the ray-pool kernel (lanes pull rays from a shared-memory queue, dmf_reverse.cuh): same results and
        // counters, 26 of 32 lanes stepping instead of 11 -- and SLOWER (2.06 vs 1.53 ms per 128 views at S512, 6.4 vs 3.8 ms at
        // S1024, profiles/r02_reverse_pool_ab.txt): the march is bound by the latency of its dependent byte loads, which the
        // one-thread-per-voxel kernel hides with 64 resident warps at 32 registers; the pool needs 64 registers (32 warps).  Kept as
        // the measured alternative, not the default.
        static const bool no_pool = std::getenv("DMF_REVERSE_POOL") == nullptr;
        if (c->reverse_format == DMF_GRID_BYTE && !no_pool) k_reverse_pool<<<dim3((unsigned)((c->n_occ + RP_BLOCK_VOX - 1) / RP_BLOCK_VOX), n_views), RP_THREADS, 0, st>>>(a);
        else if (c->reverse_format == DMF_GRID_BYTE) k_reverse<true, 1><<<grid, 128, 0, st>>>(a);
        else k_reverse<true, 0><<<grid, 128, 0, st>>>(a);
    } else {
        size_t total = (size_t)a.nax[0] * a.nax[1] * a.nax[2];
        if (!total) return 0;
        size_t gx = (total + 127) / 128;
        if (gx > 0x7fffffffull) return fail("grid too large for the whole-grid reverse scan");
        dim3 grid((unsigned)gx, n_views);
        if (c->reverse_format == DMF_GRID_BYTE) k_reverse<false, 1><<<grid, 128, 0, st>>>(a); else k_reverse<false, 0><<<grid, 128, 0, st>>>(a);
    }
    if (!c->capturing) { DMF_CUDA(cudaEventRecord(c->ev_h1, st)); c->hot_timed = true; }
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

int dmf_reverse_dev(dmf_ctx* c, int fast, int viz, const float* d_poses, int n_views, const dmf_reverse_out* d_out, void* stream) {
    if (!c || !d_out) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    if (d_out->ids || d_out->ids_offsets) return fail("dmf_reverse_dev does not produce id lists; use dmf_reverse or the visibility bitset");
    cudaStream_t st = pick_stream(c, stream);
    DMF_TRY(order_after_previous(c, st));
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    DMF_TRY(enqueue_reverse(c, fast, viz, d_poses, n_views, (unsigned*)d_out->visibility, (unsigned*)d_out->unoccluded, d_out->found_any, nullptr, nullptr, 0, st));
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    return mark_last(c, st);
}

int dmf_reverse(dmf_ctx* c, int fast, int viz, const float* poses, int n_views, const dmf_reverse_out* out) {
    if (!c || !out || (!poses && n_views > 0)) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    if (!c->vol_set) return fail("no volume uploaded");
    const size_t vw = (c->n_occ + 63) / 64;
    const bool want_ids = out->ids_offsets != nullptr;
    if (want_ids) out->ids_offsets[0] = 0;
    cudaStream_t st = c->stream;
    DMF_TRY(order_after_previous(c, st));
    const unsigned emit_cap = (unsigned)std::min<size_t>(2 * c->n_occ + 64, 0x7fffffffu);
    size_t per_view = 48 + 2 * vw * 8 + 16 + (want_ids ? (fast ? c->n_occ * 4 : (size_t)emit_cap * 16) : 0);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)n_views, (size_t)4096, ((size_t)1 << 30) / per_view}));
    int64_t ids_total = 0;
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    for (int v0 = 0; v0 < n_views; v0 += chunk) {
        const int nv = std::min(chunk, n_views - v0);
        DevBuf* ob = c->d_out[0];
        DMF_TRY(c->d_poses[0].reserve((size_t)nv * 48));
        DMF_CUDA(cudaMemcpyAsync(c->d_poses[0].p, poses + 12 * (size_t)v0, (size_t)nv * 48, cudaMemcpyHostToDevice, st));
        DMF_TRY(ob[3].reserve(std::max<size_t>(nv * vw * 8, 8))); DMF_TRY(ob[6].reserve(std::max<size_t>(nv * vw * 8, 8))); DMF_TRY(ob[4].reserve((size_t)nv * 4));
        unsigned* d_vis = ob[3].as<unsigned>(); unsigned* d_unocc = ob[6].as<unsigned>(); int* d_found = ob[4].as<int>();
        u64* d_emit = nullptr; unsigned* d_emit_count = nullptr;
        if (want_ids && !fast) {
            DMF_TRY(c->d_misc[0].reserve((size_t)nv * emit_cap * 16)); DMF_TRY(c->d_misc[1].reserve((size_t)nv * 4));
            d_emit = c->d_misc[0].as<u64>(); d_emit_count = c->d_misc[1].as<unsigned>();
        }
        const bool ids_fast = want_ids && fast;
        if (!ids_fast) DMF_TRY(enqueue_reverse(c, fast, viz, c->d_poses[0].as<float>(), nv, d_vis, d_unocc, d_found, d_emit, d_emit_count, emit_cap, st));
        bool found_delivered = false;
        long long cap_dev = 0; size_t off_bytes = 0, found_bytes = 0, first_ids = 0; char* hs = nullptr;
        if (ids_fast) {
            // emission order of reverseRayTraceFast == occupied order: expand each view's bitset in ascending order.  Counts ->
            // offsets -> gather stay on the device; one copy of (offsets, found flags, first ids) into pinned staging, one sync.
            cap_dev = (long long)nv * (long long)std::max<size_t>(c->n_occ, 1);
            off_bytes = (size_t)(nv + 1) * 8; found_bytes = ((size_t)nv * 4 + 7) / 8 * 8;
            first_ids = (size_t)std::min<long long>(cap_dev, 128 * 1024);
            DMF_TRY(c->stage.reserve(off_bytes + found_bytes + first_ids * 8));
            DMF_TRY(c->d_offsets.reserve((size_t)(nv + 1) * 8)); DMF_TRY(c->d_n_ids.reserve((size_t)nv * 4));
            hs = (char*)c->stage.p;
            const int nbb = (int)(((c->n_occ + 31) / 32 + BITS_TILE - 1) / BITS_TILE);
            if (c->n_occ) {
                DMF_TRY(c->d_out_occ.reserve((size_t)nv * c->n_occ * 4)); DMF_TRY(c->d_ids.reserve((size_t)cap_dev * 8));
                DMF_TRY(c->d_misc[2].reserve((size_t)nv * nbb * 4)); DMF_TRY(c->d_misc[3].reserve(std::max<size_t>((size_t)nv * nbb * 4, 64)));
            }
            auto enqueue_all = [&]() -> int {
                DMF_TRY(enqueue_reverse(c, fast, viz, c->d_poses[0].as<float>(), nv, d_vis, d_unocc, d_found, d_emit, d_emit_count, emit_cap, st));
                if (c->n_occ) {
                    k_bits_count<<<dim3(nbb, nv), BITS_TILE, 0, st>>>(d_vis, (int)(vw * 2), (int)c->n_occ, c->d_misc[2].as<unsigned>(), nbb);
                    k_win_offsets<<<nv, 1024, 0, st>>>(c->d_misc[2].as<unsigned>(), c->d_misc[3].as<unsigned>(), c->d_n_ids.as<int>(), nbb);
                    k_bits_emit<<<dim3(nbb, nv), BITS_TILE, 0, st>>>(d_vis, (int)(vw * 2), (int)c->n_occ, c->d_misc[3].as<unsigned>(), nbb, c->d_out_occ.as<int>(), (int)c->n_occ);
                    k_gather_ids<<<dim3(32, nv), 256, 0, st>>>(c->d_out_occ.as<int>(), c->d_offsets.as<long long>(), c->d_centroid_hash.as<u64>(), c->d_ids.as<u64>(), (int)c->n_occ, cap_dev, c->d_n_ids.as<int>());
                    c->launches += 4;
                    DMF_CUDA(cudaGetLastError());
                    DMF_CUDA(cudaMemcpyAsync(hs, c->d_offsets.p, off_bytes, cudaMemcpyDeviceToHost, st));
                    DMF_CUDA(cudaMemcpyAsync(hs + off_bytes + found_bytes, c->d_ids.p, first_ids * 8, cudaMemcpyDeviceToHost, st));
                }
                DMF_CUDA(cudaMemcpyAsync(hs + off_bytes, d_found, (size_t)nv * 4, cudaMemcpyDeviceToHost, st));
                return 0;
            };
            if (n_views == 1 && !viz && !(out->visibility && vw) && !(out->unoccluded && vw)) {
                // the shape of the drop-in's reverseRayTraceFast(volume, T, false): one captured graph from the second identical call on
                struct { int H, W, fmt, bytes_built; float K[9]; unsigned long long n_occ, epoch; } k;
                std::memset(&k, 0, sizeof k);
                k.H = c->H; k.W = c->W; k.fmt = c->reverse_format; k.bytes_built = c->bytes_built ? 1 : 0; std::memcpy(k.K, c->K, sizeof k.K); k.n_occ = c->n_occ; k.epoch = c->volume_epoch;
                DMF_TRY(run_maybe_graphed(c, c->graph_rev_ids, fnv1a(&k, sizeof k), st, enqueue_all));
            } else DMF_TRY(enqueue_all());
            DMF_CUDA(cudaStreamSynchronize(st));
            if (!c->n_occ) std::memset(hs, 0, off_bytes);
            const long long* offs = (const long long*)hs;
            const long long total = offs[nv];
            if ((size_t)(ids_total + total) > out->ids_capacity || (!out->ids && total > 0)) return fail("ids_capacity %zu too small (need >= %lld)", out->ids_capacity, (long long)(ids_total + total));
            if (total > 0) {
                const size_t head = (size_t)std::min<long long>(total, (long long)first_ids);
                std::memcpy(out->ids + ids_total, hs + off_bytes + found_bytes, head * 8);
                if ((size_t)total > head) DMF_CUDA(cudaMemcpy(out->ids + ids_total + head, c->d_ids.as<u64>() + head, ((size_t)total - head) * 8, cudaMemcpyDeviceToHost));
            }
            for (int i = 0; i < nv; i++) out->ids_offsets[v0 + i + 1] = ids_total + offs[i + 1];
            if (out->found_any) { std::memcpy(out->found_any + v0, hs + off_bytes, (size_t)nv * 4); found_delivered = true; }
            ids_total += total;
        }
        if (want_ids && !fast) {
            // whole-grid scan: the kernel appended (scan index, centroid hash) pairs in arbitrary order; the reference's
            // push_back order is the scan order, so sort each view's (short) list by scan index.
            std::vector<unsigned> cnt(nv);
            DMF_CUDA(cudaMemcpyAsync(cnt.data(), d_emit_count, (size_t)nv * 4, cudaMemcpyDeviceToHost, st));
            DMF_CUDA(cudaStreamSynchronize(st));
            for (int i = 0; i < nv; i++) {
                if (cnt[i] > emit_cap) return fail("reverseRayTrace emitted %u ids for view %d, more than the %u slots reserved", cnt[i], v0 + i, emit_cap);
                if ((size_t)(ids_total + cnt[i]) > out->ids_capacity || (!out->ids && cnt[i] > 0)) return fail("ids_capacity %zu too small", out->ids_capacity);
                std::vector<uint64_t> pairs(2 * (size_t)cnt[i]);
                if (cnt[i]) DMF_CUDA(cudaMemcpy(pairs.data(), d_emit + 2 * (size_t)i * emit_cap, pairs.size() * 8, cudaMemcpyDeviceToHost));
                std::vector<std::pair<uint64_t, uint64_t>> pv(cnt[i]);
                for (unsigned j = 0; j < cnt[i]; j++) pv[j] = {pairs[2 * j], pairs[2 * j + 1]};
                std::sort(pv.begin(), pv.end());
                for (unsigned j = 0; j < cnt[i]; j++) out->ids[ids_total + j] = pv[j].second;
                ids_total += cnt[i];
                out->ids_offsets[v0 + i + 1] = ids_total;
            }
        }
        if (out->visibility && vw) DMF_CUDA(cudaMemcpyAsync(out->visibility + v0 * vw, d_vis, nv * vw * 8, cudaMemcpyDeviceToHost, st));
        if (out->unoccluded && vw) DMF_CUDA(cudaMemcpyAsync(out->unoccluded + v0 * vw, d_unocc, nv * vw * 8, cudaMemcpyDeviceToHost, st));
        if (out->found_any && !found_delivered) DMF_CUDA(cudaMemcpyAsync(out->found_any + v0, d_found, (size_t)nv * 4, cudaMemcpyDeviceToHost, st));
        // (the fast id path has already synchronised; another sync is needed only if something was enqueued after it)
        if ((out->visibility && vw) || (out->unoccluded && vw) || (out->found_any && !found_delivered) || !(want_ids && fast)) DMF_CUDA(cudaStreamSynchronize(st));
    }
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    return 0;
}

int dmf_zbuffer(dmf_ctx* c, const float pose[12], int32_t* depth, int64_t* n_splat) {
    if (!c || !pose) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    RevArgs a; DMF_TRY(fill_rev_args(c, a));
    cudaStream_t st = c->stream;
    const size_t HW = (size_t)c->H * c->W;
    DMF_TRY(c->d_poses[0].reserve(48)); DMF_TRY(c->d_inv_poses.reserve(48)); DMF_TRY(c->d_misc[2].reserve(HW * 4)); DMF_TRY(c->d_misc[3].reserve(16));
    DMF_CUDA(cudaMemcpyAsync(c->d_poses[0].p, pose, 48, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    k_invert_poses<<<1, 32, 0, st>>>(c->d_poses[0].as<float>(), c->d_inv_poses.as<float>(), 1);
    DMF_TRY(fill_u32(c, st, c->d_misc[2].p, HW, 0x7fffffffu));
    DMF_CUDA(cudaMemsetAsync(c->d_misc[3].p, 0, 16, st));
    a.poses = c->d_poses[0].as<float>(); a.inv_poses = c->d_inv_poses.as<float>(); a.zbuf = c->d_misc[2].as<int>();
    const size_t total = (size_t)a.nax[0] * a.nax[1] * a.nax[2];
    unsigned long long* cnt = c->d_misc[3].as<unsigned long long>();
    if (total) {
        k_zbuffer<0><<<blocks_for(total, 256, 148 * 32), 256, 0, st>>>(a, cnt, cnt + 1);
        k_zbuffer<1><<<blocks_for(total, 256, 148 * 32), 256, 0, st>>>(a, cnt, cnt + 1);
    }
    k_finish_zbuf<<<blocks_for(HW, 256), 256, 0, st>>>(a.zbuf, HW);
    c->launches += 4;
    DMF_CUDA(cudaGetLastError());
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    unsigned long long h_cnt[2] = {0, 0};
    DMF_CUDA(cudaMemcpyAsync(h_cnt, cnt, 16, cudaMemcpyDeviceToHost, st));
    if (depth) DMF_CUDA(cudaMemcpyAsync(depth, a.zbuf, HW * 4, cudaMemcpyDeviceToHost, st));
    DMF_CUDA(cudaStreamSynchronize(st));
    if (n_splat) *n_splat = (int64_t)h_cnt[0];
    return 0;
}

// setDimensions + setVolumeSize + constructVolume + integratePointCloud(cloud, normals) with the integration on the GPU
// (dmf_integrate.cuh) and the march structures built from its output where it lies (dmf_volume.cuh): the only host
// traffic is the point cloud in and two counts out.  Same result as dmf_volume_from_points: occupied_cells_ in
// first-insertion order, normals per voxel in point order.
int dmf_volume_from_points_gpu(dmf_ctx* c, const double bounds[6], const int dims[3], const float* xyz, const float* normals, size_t n) {
    if (!c || !bounds || !dims || (!xyz && n > 0)) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_CUDA(cudaDeviceSynchronize());
    if (n >= 0xFFFFFFFFull) return fail("too many points for 32-bit point indices");
    for (int a = 0; a < 3; a++) if (dims[a] < 1) return fail("bad dims");
    HostVolume hv;
    hv.construct(bounds, dims);
    DMF_TRY(set_volume_geometry(c, bounds, hv.delta, hv.dim));
    c->vol_set = false; c->mirror_valid = false;
    cudaStream_t st = c->stream;
    const VolDev v = c->vol;                                       // scalars only are used by the K0 kernels
    const size_t ncell = (size_t)v.pdim[0] * v.pdim[1] * v.pdim[2];
    DevBuf d_xyz, d_nrm, d_key, d_first, d_flag, d_ord, d_cnt, d_cur, d_pidx;
    auto cleanup = [&]() { for (DevBuf* b : {&d_xyz, &d_nrm, &d_key, &d_first, &d_flag, &d_ord, &d_cnt, &d_cur, &d_pidx}) b->release(); };
    int rc = 0;
    unsigned n_occ = 0, n_nrm = 0;
    do {
        if ((rc = c->d_err.reserve(16))) break;
        if (n == 0) break;
        if ((rc = d_xyz.reserve(n * 12)) || (rc = d_key.reserve(n * 4)) || (rc = d_first.reserve(ncell * 4)) || (rc = d_flag.reserve(n * 4)) || (rc = d_ord.reserve(n * 4))) break;
        if (cudaMemcpyAsync(d_xyz.p, xyz, n * 12, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = fail("H2D of the points failed"); break; }
        const unsigned g = blocks_for(n, 256, 148 * 16);
        k_pt_key<<<g, 256, 0, st>>>(v, d_xyz.as<float>(), n, d_key.as<unsigned>());
        if ((rc = fill_u32(c, st, d_first.p, ncell, 0xFFFFFFFFu))) break;
        k_pt_first<<<g, 256, 0, st>>>(d_key.as<unsigned>(), n, d_first.as<unsigned>());
        k_pt_flag<<<g, 256, 0, st>>>(d_key.as<unsigned>(), d_first.as<unsigned>(), n, d_flag.as<unsigned>());
        c->launches += 3;
        if ((rc = scan_u32_async(c, st, d_flag.as<unsigned>(), d_ord.as<unsigned>(), n, c->d_err.as<unsigned>()))) break;
        if (cudaMemcpyAsync(&n_occ, c->d_err.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = fail("device integratePointCloud failed (occupied count)"); break; }
        if (n_occ == 0) break;
        if ((rc = c->d_occ_ids.reserve((size_t)n_occ * 8))) break;
        k_pt_assign<<<g, 256, 0, st>>>(v, d_key.as<unsigned>(), d_flag.as<unsigned>(), d_ord.as<unsigned>(), n, d_first.as<unsigned>(), c->d_occ_ids.as<u64>());
        c->launches++;
        if (normals) {
            if ((rc = d_nrm.reserve(n * 12)) || (rc = d_cnt.reserve(((size_t)n_occ + 1) * 4)) || (rc = c->d_noff.reserve(((size_t)n_occ + 1) * 4)) || (rc = d_cur.reserve((size_t)n_occ * 4))) break;
            if (cudaMemcpyAsync(d_nrm.p, normals, n * 12, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = fail("H2D of the normals failed"); break; }
            cudaMemsetAsync(d_cnt.p, 0, ((size_t)n_occ + 1) * 4, st); cudaMemsetAsync(d_cur.p, 0, (size_t)n_occ * 4, st);
            k_pt_count<<<g, 256, 0, st>>>(d_key.as<unsigned>(), d_first.as<unsigned>(), n, d_cnt.as<unsigned>());
            c->launches++;
            if ((rc = scan_u32_async(c, st, d_cnt.as<unsigned>(), c->d_noff.as<unsigned>(), (size_t)n_occ + 1, c->d_err.as<unsigned>()))) break;
            if (cudaMemcpyAsync(&n_nrm, c->d_err.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { rc = fail("device integratePointCloud failed (normal count)"); break; }
            if ((rc = d_pidx.reserve(std::max<size_t>(n_nrm, 1) * 4)) || (rc = c->d_normals.reserve(std::max<size_t>(n_nrm, 1) * 12))) break;
            k_pt_scatter<<<g, 256, 0, st>>>(d_key.as<unsigned>(), d_first.as<unsigned>(), c->d_noff.as<unsigned>(), n, d_cur.as<unsigned>(), d_pidx.as<unsigned>());
            k_seg_sort<<<(n_occ + 127) / 128, 128, 0, st>>>(c->d_noff.as<unsigned>(), n_occ, d_pidx.as<unsigned>());
            k_gather_normals<<<blocks_for(n_nrm, 256, 148 * 16), 256, 0, st>>>(d_pidx.as<unsigned>(), d_nrm.as<float>(), n_nrm, c->d_normals.as<float>());
            c->launches += 3;
        }
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = fail("device integratePointCloud failed: %s", cudaGetErrorString(e)); break; }
    } while (0);
    cleanup();
    if (rc) return rc;
    if (!normals || n_occ == 0) {                                   // no normal lists: an all-zero CSR
        if ((rc = c->d_noff.reserve(((size_t)n_occ + 1) * 4))) return rc;
        DMF_CUDA(cudaMemsetAsync(c->d_noff.p, 0, ((size_t)n_occ + 1) * 4, st));
        n_nrm = 0;
    }
    return build_volume_device(c, n_occ, n_nrm);
}

int dmf_segments_collide(dmf_ctx* c, const float* a, const float* b, int n, int guard_coords, uint8_t* out) {
    if (!c || (!a && n > 0) || (!b && n > 0) || (!out && n > 0)) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    if (!c->vol_set) return fail("no volume uploaded");
    if (n <= 0) return 0;
    cudaStream_t st = c->stream;
    if (c->reverse_format == DMF_GRID_BYTE) DMF_TRY(ensure_bytes(c, st));
    DMF_TRY(c->d_misc[0].reserve((size_t)n * 12)); DMF_TRY(c->d_misc[1].reserve((size_t)n * 12)); DMF_TRY(c->d_misc[2].reserve((size_t)n));
    DMF_CUDA(cudaMemcpyAsync(c->d_misc[0].p, a, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaMemcpyAsync(c->d_misc[1].p, b, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    DMF_CUDA(cudaEventRecord(c->ev_h0, st));
    if (c->reverse_format == DMF_GRID_BYTE)
        k_segments_collide<1><<<(n + 127) / 128, 128, 0, st>>>(c->vol, c->d_misc[0].as<float>(), c->d_misc[1].as<float>(), n, guard_coords, c->d_misc[2].as<unsigned char>(), c->d_counters.as<u64>());
    else
        k_segments_collide<0><<<(n + 127) / 128, 128, 0, st>>>(c->vol, c->d_misc[0].as<float>(), c->d_misc[1].as<float>(), n, guard_coords, c->d_misc[2].as<unsigned char>(), c->d_counters.as<u64>());
    DMF_CUDA(cudaEventRecord(c->ev_h1, st));
    c->hot_timed = true;
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    DMF_CUDA(cudaMemcpyAsync(out, c->d_misc[2].p, (size_t)n, cudaMemcpyDeviceToHost, st));
    DMF_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Algorithms::moveCamera / repositionCamera (Algorithms.hpp:170-188) in float: t -= row2(linear) * distance / 1000
static void reposition_camera(const float* pose12, unsigned distance, float* out12) {
    std::memcpy(out12, pose12, 48);
    const float d = (float)(double)distance;
    for (int i = 0; i < 3; i++) {
        volatile float prod = pose12[8 + i] * d;       // separate roundings, as Eigen evaluates n*distance/1000.0
        volatile float q = prod / 1000.0f;
        out12[4 * i + 3] = pose12[4 * i + 3] - q;
    }
}

int dmf_optimize_standoff(dmf_ctx* c, const float* poses, int n, unsigned low0, unsigned high0, uint32_t* mid_out, float* poses_out) {
    if (!c || (!poses && n > 0)) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    if (!c->vol_set) return fail("no volume uploaded");
    if (n <= 0) return 0;
    std::vector<unsigned> low(n, low0), high(n, high0), mid(n, low0);
    cudaStream_t st = c->stream;
    for (;;) {
        std::vector<int> act;
        for (int i = 0; i < n; i++) if (low[i] < high[i]) act.push_back(i);
        if (act.empty()) break;
        const int nv = 2 * (int)act.size();
        std::vector<float> batch(12 * (size_t)nv);
        for (size_t j = 0; j < act.size(); j++) {
            reposition_camera(poses + 12 * (size_t)act[j], low[act[j]], &batch[12 * (2 * j)]);
            reposition_camera(poses + 12 * (size_t)act[j], high[act[j]], &batch[12 * (2 * j + 1)]);
        }
        std::vector<unsigned> counts(nv, 0);
        for (int v0 = 0; v0 < nv; v0 += 4096) {                       // reverseRayTrace launches one thread per scan position and view
            const int m = std::min(4096, nv - v0);
            DMF_TRY(c->d_poses[0].reserve((size_t)m * 48)); DMF_TRY(c->d_misc[1].reserve((size_t)m * 4));
            DMF_CUDA(cudaMemcpyAsync(c->d_poses[0].p, &batch[12 * (size_t)v0], (size_t)m * 48, cudaMemcpyHostToDevice, st));
            DMF_TRY(enqueue_reverse(c, /*fast=*/0, /*viz=*/0, c->d_poses[0].as<float>(), m, nullptr, nullptr, nullptr, nullptr, c->d_misc[1].as<unsigned>(), 0, st));
            DMF_CUDA(cudaMemcpyAsync(&counts[v0], c->d_misc[1].p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
            DMF_CUDA(cudaStreamSynchronize(st));
        }
        for (size_t j = 0; j < act.size(); j++) {                     // Algorithms.hpp:409-417
            const int i = act[j];
            mid[i] = (low[i] + high[i]) / 2;
            if (counts[2 * j + 1] > counts[2 * j]) low[i] = mid[i] + 1; else high[i] = mid[i];
        }
    }
    for (int i = 0; i < n; i++) {
        if (mid_out) mid_out[i] = mid[i];
        if (poses_out) reposition_camera(poses + 12 * (size_t)i, mid[i], poses_out + 12 * (size_t)i);
    }
    return 0;
}

int dmf_set_reverse_format(dmf_ctx* c, int grid_format) {
    if (!c) return fail("null context");
    if (grid_format != DMF_GRID_BIT && grid_format != DMF_GRID_BYTE) return fail("bad grid_format %d", grid_format);
    c->reverse_format = grid_format;
    return 0;
}

int dmf_selftest_div1000(dmf_ctx* c, uint64_t mismatches[5]) {
    if (!c || !mismatches) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    DMF_TRY(c->d_misc[3].reserve(40));
    const unsigned long long init[5] = {0, 0, 0, 0xffffffffull, 0};
    DMF_CUDA(cudaMemcpyAsync(c->d_misc[3].p, init, 40, cudaMemcpyHostToDevice, c->stream));
    k_selftest_div1000<<<148 * 16, 256, 0, c->stream>>>(c->d_misc[3].as<unsigned long long>());
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    DMF_CUDA(cudaMemcpyAsync(mismatches, c->d_misc[3].p, 40, cudaMemcpyDeviceToHost, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int dmf_or_reduce_dev(dmf_ctx* c, uint64_t* d_dst, const uint64_t* d_src, int n_src, size_t words, void* stream) {
    if (!c) return fail("null context");
    DMF_CUDA(cudaSetDevice(c->device));
    if (!words || n_src <= 0) return 0;
    k_or_reduce<<<blocks_for(words, 256, 148 * 8), 256, 0, pick_stream(c, stream)>>>((u64*)d_dst, (const u64*)d_src, n_src, words);
    c->launches++;
    DMF_CUDA(cudaGetLastError());
    return 0;
}

static int greedy_set_cover_strided(dmf_ctx* c, const uint64_t* d_bits, int n_sets, size_t words, size_t stride, int32_t* selected, int* n_selected);
int dmf_greedy_set_cover_dev(dmf_ctx* c, const uint64_t* d_bits, int n_sets, size_t words, int32_t* selected, int* n_selected) {
    return greedy_set_cover_strided(c, d_bits, n_sets, words, words, selected, n_selected);
}
static int greedy_set_cover_strided(dmf_ctx* c, const uint64_t* d_bits, int n_sets, size_t words, size_t stride, int32_t* selected, int* n_selected) {
    if (!c || !selected || !n_selected) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    *n_selected = 0;
    if (n_sets <= 0 || !words) return 0;
    cudaStream_t st = c->stream;
    DevBuf &covered = c->d_misc[0], &taken = c->d_misc[1], &gain = c->d_misc[2], &result = c->d_misc[3];
    DMF_TRY(covered.reserve(words * 8)); DMF_TRY(taken.reserve((size_t)n_sets * 4)); DMF_TRY(gain.reserve((size_t)n_sets * 4)); DMF_TRY(result.reserve(16));
    DMF_CUDA(cudaMemsetAsync(covered.p, 0, words * 8, st));
    DMF_CUDA(cudaMemsetAsync(taken.p, 0, (size_t)n_sets * 4, st));
    DMF_CUDA(cudaEventRecord(c->ev_k0, st));
    while (true) {
        k_cover_gain<<<n_sets, 256, 0, st>>>((const u64*)d_bits, covered.as<u64>(), taken.as<int>(), words, stride, gain.as<unsigned>());
        k_cover_pick<<<1, 1024, 0, st>>>(gain.as<unsigned>(), n_sets, result.as<int>());
        k_cover_apply<<<blocks_for(words, 256, 148 * 4), 256, 0, st>>>((const u64*)d_bits, covered.as<u64>(), taken.as<int>(), words, stride, result.as<int>());
        c->launches += 3;
        DMF_CUDA(cudaGetLastError());
        int res[2];
        DMF_CUDA(cudaMemcpyAsync(res, result.p, 8, cudaMemcpyDeviceToHost, st));
        DMF_CUDA(cudaStreamSynchronize(st));
        if (res[0] < 0) break;       // selected == -1  (Algorithms.hpp:71)
        if (res[1] < 5) break;       // max_points < 5  (Algorithms.hpp:73)
        selected[(*n_selected)++] = res[0];
        if (*n_selected >= n_sets) break;
    }
    DMF_CUDA(cudaEventRecord(c->ev_k1, st));
    c->timed = true;
    return 0;
}

int dmf_greedy_set_cover(dmf_ctx* c, const uint64_t* bitsets, int n_sets, size_t words, int32_t* selected, int* n_selected) {
    if (!c || !bitsets) return fail("null argument");
    DMF_CUDA(cudaSetDevice(c->device));
    if (n_sets <= 0 || !words) { if (n_selected) *n_selected = 0; return 0; }
    DMF_TRY(c->d_ids.reserve((size_t)n_sets * words * 8));
    DMF_CUDA(cudaMemcpyAsync(c->d_ids.p, bitsets, (size_t)n_sets * words * 8, cudaMemcpyHostToDevice, c->stream));
    return dmf_greedy_set_cover_dev(c, c->d_ids.as<uint64_t>(), n_sets, words, selected, n_selected);
}

}  // extern "C"
