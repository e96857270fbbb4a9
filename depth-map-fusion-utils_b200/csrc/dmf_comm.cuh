// dmf_comm.cuh -- the view sweep over the GPUs of one box, inside the library (include/dmf_b200.h "multi-GPU").
// Included at the end of dmf_b200.cu (same translation unit, same helpers).
//
// The reference's drivers loop over candidate views on one CPU thread (tests/SetCover.cpp:218-240,
// tests/CameraMotionPlanner.cpp:334-356, tests/CameraPathGen.cpp:158-180).  Views are independent, so they are dealt
// round-robin over the GPUs (rank r marches views r, r+N, ...: neighbouring views cost about the same, so every GPU gets
// the same mix), with the volume replicated on every GPU.  What has to cross GPUs:
//   * every view's visibility row, to every GPU (the set-cover consumer).  Rows are disjoint by view, so no collective is
//     needed: the march writes its rows in place in its own copy of the gathered [n_views][row_words] buffer, and the finished
//     rows are stored straight into every peer's copy through peer-mapped pointers (16-byte stores over NVLink); a sequence
//     flag raised in every peer (system-scope release) tells consumers, which wait on their own flag words
//     (cuStreamWaitValue32), that this GPU's rows have landed.  Two issuers of those stores exist: a small push kernel right
//     behind the march (k_publish_rows, the default) and the march kernels' own epilogue (publish_view_row in dmf_device.cuh:
//     a ticket per view finds the last block, DMF_COMM_EXCHANGE=epilogue).  Measured on B200 the epilogue variant costs the
//     march ~20 % (barrier + fence + atomic round trip at the tail of every ~8 us block) to overlap a transfer that takes a
//     few microseconds, so the push kernel is the default.  Where peer mapping is not possible the rows go through
//     ncclAllGather + an interleave kernel instead (DMF_COMM_EXCHANGE=nccl forces that path: the A/B baseline);
//   * the observed (occupied/free) grids and the Voxel::view / Voxel::good marks of a sharded fusion run: bitwise OR, resp.
//     min over the first view id, as a reduce-scatter + all-gather over peer memory: every GPU reduces 1/N of the words from
//     all peers and pushes the result back (k_peer_reduce), so each link carries 2 (N-1)/N of a grid instead of N-1 grids.
// Two ways to form a group: dmf_comm_init_all (ONE process drives all GPUs -- the shape the reference's single-threaded C++
// drivers need; peers are addressed directly after cudaDeviceEnablePeerAccess) and dmf_comm_init_rank (one process per GPU,
// torchrun / MPI style; arenas are mapped with CUDA IPC, handles travel through ncclAllGather).
// NCCL is resolved with dlopen at run time, so the library still loads (single GPU) where NCCL is absent.
#pragma once
#include <dlfcn.h>

namespace {

// ---- NCCL, resolved at run time (types restated: only pointers, one 128-byte id and small enums cross the boundary) --------
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string why;
};
constexpr int kNcclUint8 = 1, kNcclInt32 = 2, kNcclMin = 3;

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    // by soname: a process that already carries NCCL (PyTorch) shares its copy
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { api.why = dlerror() ? dlerror() : "libnccl.so.2 not found"; return nullptr; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) { ok = false; api.why = std::string("missing symbol ") + n; } return p; };
    api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
    api.CommInitAll = (int (*)(NcclComm*, int, const int*))sym("ncclCommInitAll");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclBroadcast");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    if (!ok) { dlclose(api.lib); api.lib = nullptr; return nullptr; }
    return &api;
}
#define DMF_NCCL(call) do { int r_ = (call); if (r_ != 0) return dmf::fail("%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); } while (0)

// ---- kernels of the exchange ---------------------------------------------------------------------------------------------
struct SignalArgs { unsigned* flag[DMF_MAX_PEERS]; int n; unsigned value; };
// raise this rank's flag word in every peer (after everything this stream did before is visible system-wide)
__global__ void k_signal(const SignalArgs a) {
    __threadfence_system();
    if ((int)threadIdx.x < a.n) *reinterpret_cast<volatile unsigned*>(a.flag[threadIdx.x]) = a.value;
}
// fallback for cuStreamWaitValue32: one thread polls this GPU's own flag words
__global__ void k_wait_flags(const unsigned* flags, const int* ranks, int n, unsigned value) {
    if ((int)threadIdx.x < n) {
        const volatile unsigned* f = flags + ranks[threadIdx.x];
        while ((int)(*f - value) < 0) __nanosleep(200);
    }
    __threadfence_system();
}

// Default exchange: ONE small kernel behind the march pushes this GPU's finished rows into every peer's gathered buffer
// (16-byte peer stores over NVLink), writes each row's found word, and the block that finishes last raises this GPU's flag
// in every peer.  Measured against the in-kernel epilogue (publish_view_row) it wins: the ticket protocol adds a barrier, a
// fence and an atomic round trip to the tail of every 8-microsecond march block (+20 % on k_forward_line), while the rows of
// a whole step are 2 MB -- a few microseconds of NVLink time that need no overlap.
__global__ void __launch_bounds__(256) k_publish_rows(const PubTable pub, u64* __restrict__ own_rows, size_t pitch_words) {
    const int j = blockIdx.x;                                  // local view
    u64* const row = own_rows + (size_t)j * pitch_words;
    const unsigned found = pub.found_any ? (unsigned)pub.found_any[j] : 0u;
    const size_t g_off = (size_t)(pub.row0 + j * pub.row_step) * pub.row_words;
    const unsigned n16 = pub.row_words >> 1;
    for (unsigned i = threadIdx.x; i < n16; i += blockDim.x) {
        uint4 val = reinterpret_cast<const uint4*>(row)[i];
        if (2u * i <= pub.vis_words && pub.vis_words < 2u * i + 2u) {           // the piece that holds the found word
            if (pub.vis_words & 1u) { val.z = found; val.w = 0u; } else { val.x = found; val.y = 0u; }
            reinterpret_cast<uint4*>(row)[i] = val;
        }
        for (int p = 0; p < pub.n_peers; p++) reinterpret_cast<uint4*>(pub.peer_rows[p] + g_off)[i] = val;
    }
    if (pub.n_peers == 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(pub.views_done, 1u) == (unsigned)pub.n_views_pass - 1u) {
        __threadfence_system();
        for (int p = 0; p < pub.n_peers; p++) *reinterpret_cast<volatile unsigned*>(pub.peer_flag[p]) = pub.seq;
    }
}

struct PeerReduceArgs { unsigned* stage[DMF_MAX_PEERS]; int world; size_t lo, hi; };     // [lo, hi) in 16-byte units: this rank's slice
// OP 0: bitwise OR, OP 1: min over int32.  Reads the slice from every member's staging copy (peer loads), pushes the result
// into every member's staging copy (peer stores).  Slices are disjoint by rank, so nobody reads what another rank writes.
template <int OP>
__global__ void __launch_bounds__(256) k_peer_reduce(const PeerReduceArgs a) {
    for (size_t i = a.lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.hi; i += (size_t)gridDim.x * blockDim.x) {
        uint4 acc = reinterpret_cast<const uint4*>(a.stage[0])[i];
        for (int r = 1; r < a.world; r++) {
            const uint4 v = reinterpret_cast<const uint4*>(a.stage[r])[i];
            if (OP == 0) { acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w; }
            else { acc.x = (unsigned)min((int)acc.x, (int)v.x); acc.y = (unsigned)min((int)acc.y, (int)v.y); acc.z = (unsigned)min((int)acc.z, (int)v.z); acc.w = (unsigned)min((int)acc.w, (int)v.w); }
        }
        for (int r = 0; r < a.world; r++) reinterpret_cast<uint4*>(a.stage[r])[i] = acc;
    }
}
// NCCL path: rows gathered rank-major [world][n_max][row_words] -> view order (view g = rank + j * world)
__global__ void k_interleave_rows(const u64* __restrict__ gathered, u64* __restrict__ rows, int world, int n_max, int n_views, unsigned row_words) {
    const size_t total = (size_t)n_views * row_words;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = i / row_words, w = i % row_words;
        rows[i] = gathered[((g % world) * (size_t)n_max + g / world) * row_words + w];
    }
}
__global__ void k_found_into_rows(u64* __restrict__ rows, const int* __restrict__ found, int n, unsigned row_words, unsigned vis_words) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) rows[(size_t)j * row_words + vis_words] = (u64)(unsigned)found[j];
}
// sharded CLASSIFY: batch-local first-view index -> global view index (INT_MAX = not hit stays)
__global__ void k_first_view_to_global(int* first_view, int n_occ, int rank, int world) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_occ && first_view[i] != 0x7fffffff) first_view[i] = rank + first_view[i] * world;
}

constexpr size_t kFlagBytes = 4096;              // [0,256): gather flags, [256,512): reduce "staged", [512,768): reduce "done"; one u32 per source rank
constexpr size_t kFlagGather = 0, kFlagStaged = 256, kFlagReduced = 512;

struct CommMember {
    dmf_ctx* ctx = nullptr;
    int rank = 0;                                // global rank
    NcclComm nccl = nullptr;
    char* arena = nullptr;                       // [flags][gather buffer 0][gather buffer 1][reduce staging]
    std::vector<char*> peer;                     // [world] every rank's arena as addressable from this member's device (own: arena)
    std::vector<char*> ipc_opened;               // mappings to close (multi-process)
    DevBuf d_ticket, d_poses, d_found, d_dense, d_nccl_gather, d_hdr, d_ranks;
    int n_my = 0;                                // views this member marched in the last sweep
};

typedef int (*WaitValue32Fn)(cudaStream_t, unsigned long long, unsigned, unsigned);

}  // namespace

struct dmf_comm {
    int world = 1, n_local = 1, rank0 = 0;
    bool single_process = true, owns_ctx = false;
    bool p2p = false;                            // every member can address every other member's arena
    int exchange = 0;                            // 0: peer stores by a push kernel behind the march, 1: NCCL all-gather, 2: peer stores from the march kernels' epilogue
    std::vector<CommMember> m;
    size_t gather_cap = 0, stage_cap = 0;        // bytes per gather buffer / of the staging region (same on every rank)
    unsigned seq = 0, rseq = 0;                  // sequence numbers of the gather and reduce flag protocols
    WaitValue32Fn wait32 = nullptr;
    // last sweep
    int last_n_views = 0, last_parity = 0;
    size_t last_row_words = 0, last_vis_words = 0;
    bool have_sweep = false;
};

namespace {

inline size_t off_gather(const dmf_comm* g, int parity) { return kFlagBytes + (size_t)parity * g->gather_cap; }
inline size_t off_stage(const dmf_comm* g) { return kFlagBytes + 2 * g->gather_cap; }
inline int my_view_count(int n_views, int rank, int world) { return rank < n_views ? (n_views - rank + world - 1) / world : 0; }

int comm_free_arenas(dmf_comm* g) {
    for (auto& mm : g->m) {
        if (!mm.ctx) continue;
        cudaSetDevice(mm.ctx->device);
        cudaDeviceSynchronize();
        for (char* p : mm.ipc_opened) cudaIpcCloseMemHandle(p);
        mm.ipc_opened.clear();
    }
    for (auto& mm : g->m) {
        if (!mm.ctx) continue;
        cudaSetDevice(mm.ctx->device);
        if (mm.arena) cudaFree(mm.arena);
        mm.arena = nullptr; mm.peer.assign(g->world, nullptr);
    }
    return 0;
}

// Collective: every rank calls with the same sizes.  (Re)allocates the arenas when they are too small and re-establishes the
// peer mappings; sequence numbers restart because the new flag words are zero.
int comm_reserve(dmf_comm* g, size_t gather_bytes, size_t stage_bytes) {
    gather_bytes = (gather_bytes + 255) / 256 * 256; stage_bytes = (stage_bytes + 255) / 256 * 256;
    if (g->m[0].arena && gather_bytes <= g->gather_cap && stage_bytes <= g->stage_cap) return 0;
    // (capacities are kept as multiples of 256 so that asking again for what is there never looks like growth)
    const size_t gcap = std::max((gather_bytes + gather_bytes / 4 + 255) / 256 * 256, g->gather_cap), scap = std::max(stage_bytes, g->stage_cap);
    comm_free_arenas(g);
    g->gather_cap = gcap; g->stage_cap = scap;
    g->seq = 0; g->rseq = 0; g->have_sweep = false;
    const size_t total = kFlagBytes + 2 * gcap + scap;
    for (auto& mm : g->m) {
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, total);
        if (e != cudaSuccess) return fail("cudaMalloc(%zu) for the exchange arena failed: %s", total, cudaGetErrorString(e));
        mm.arena = (char*)p;
        DMF_CUDA(cudaMemset(mm.arena, 0, kFlagBytes));
        mm.peer.assign(g->world, nullptr);
        mm.peer[mm.rank] = mm.arena;
    }
    if (g->world == 1) return 0;
    if (g->single_process) {
        if (g->p2p) for (auto& mm : g->m) for (auto& other : g->m) mm.peer[other.rank] = other.arena;      // UVA + peer access: the raw pointer works
        return 0;
    }
    // one process per GPU: exchange CUDA IPC handles through NCCL and map every peer's arena
    CommMember& mm = g->m[0];
    NcclApi* nc = nccl_api();
    if (!nc || !mm.nccl) return fail("multi-process group without NCCL");
    DMF_CUDA(cudaSetDevice(mm.ctx->device));
    if (!g->p2p) return 0;
    cudaIpcMemHandle_t mine;
    DMF_CUDA(cudaIpcGetMemHandle(&mine, mm.arena));
    DMF_TRY(mm.d_hdr.reserve((size_t)(g->world + 1) * sizeof mine));
    char* d_all = mm.d_hdr.as<char>() + sizeof mine;
    DMF_CUDA(cudaMemcpyAsync(mm.d_hdr.p, &mine, sizeof mine, cudaMemcpyHostToDevice, mm.ctx->stream));
    DMF_NCCL(nc->AllGather(mm.d_hdr.p, d_all, sizeof mine, kNcclUint8, mm.nccl, mm.ctx->stream));
    std::vector<cudaIpcMemHandle_t> all(g->world);
    DMF_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)g->world * sizeof mine, cudaMemcpyDeviceToHost, mm.ctx->stream));
    DMF_CUDA(cudaStreamSynchronize(mm.ctx->stream));
    for (int r = 0; r < g->world; r++) {
        if (r == mm.rank) continue;
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
        mm.peer[r] = (char*)p; mm.ipc_opened.push_back((char*)p);
    }
    return 0;
}

// stream-ordered wait until every other rank's flag word (kind) in this member's arena has reached `value`
int comm_wait(dmf_comm* g, CommMember& mm, size_t kind, unsigned value, cudaStream_t st) {
    if (g->world == 1) return 0;
    DMF_CUDA(cudaSetDevice(mm.ctx->device));
    if (g->wait32) {
        for (int r = 0; r < g->world; r++) {
            if (r == mm.rank) continue;
            const int rc = g->wait32(st, (unsigned long long)(uintptr_t)(mm.arena + kind + 4 * (size_t)r), value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
            if (rc != 0) return fail("cuStreamWaitValue32 failed (CUresult %d)", rc);
        }
        return 0;
    }
    std::vector<int> ranks;
    for (int r = 0; r < g->world; r++) if (r != mm.rank) ranks.push_back(r);
    DMF_TRY(mm.d_ranks.reserve(ranks.size() * 4));
    DMF_CUDA(cudaMemcpyAsync(mm.d_ranks.p, ranks.data(), ranks.size() * 4, cudaMemcpyHostToDevice, st));
    k_wait_flags<<<1, 32, 0, st>>>((const unsigned*)(mm.arena + kind), mm.d_ranks.as<int>(), (int)ranks.size(), value);
    mm.ctx->launches++;
    DMF_CUDA(cudaGetLastError());
    return 0;
}

int comm_signal(dmf_comm* g, CommMember& mm, size_t kind, unsigned value, cudaStream_t st) {
    if (g->world == 1) return 0;
    DMF_CUDA(cudaSetDevice(mm.ctx->device));
    SignalArgs a; a.n = 0; a.value = value;
    for (int r = 0; r < g->world; r++) if (r != mm.rank) a.flag[a.n++] = (unsigned*)(mm.peer[r] + kind + 4 * (size_t)mm.rank);
    k_signal<<<1, 32, 0, st>>>(a);
    mm.ctx->launches++;
    DMF_CUDA(cudaGetLastError());
    return 0;
}

int comm_check_same_volume(dmf_comm* g) {
    for (auto& mm : g->m) {
        if (!mm.ctx->vol_set) return fail("member %d has no volume (dmf_comm_replicate_volume)", mm.rank);
        if (!mm.ctx->cam_set) return fail("member %d has no camera (dmf_comm_set_camera)", mm.rank);
        if (mm.ctx->n_occ != g->m[0].ctx->n_occ) return fail("members hold different volumes (%zu vs %zu occupied voxels)", mm.ctx->n_occ, g->m[0].ctx->n_occ);
    }
    return 0;
}

inline size_t reduce_stage_bytes(const dmf_comm* g) {
    const dmf_ctx* c = g->m[0].ctx;
    return (std::max<size_t>(c->n_grid_words, c->n_occ) * 4 + 15) / 16 * 16 + 16;
}

struct SweepSpec {
    bool reverse = false; int fast = 1, viz = 0;
    const dmf_forward_params* fwd = nullptr;
    const dmf_forward_out* const* d_out = nullptr;    // optional per-member device outputs (depth / points / hit voxels of the member's own views)
};

// Enqueue one member's share of a sweep on `st`: zero its rows and tickets, march with the publish table, then (NCCL path)
// the all-gather + interleave.  d_poses: this member's views (j-th = global view rank + j * world), already on its device.
int sweep_member_enqueue(dmf_comm* g, CommMember& mm, const SweepSpec& spec, const float* d_poses, int n_views, cudaStream_t st, const dmf_forward_out* extra) {
    dmf_ctx* c = mm.ctx;
    DMF_CUDA(cudaSetDevice(c->device));
    const int n_my = my_view_count(n_views, mm.rank, g->world);
    mm.n_my = n_my;
    const size_t rw = g->last_row_words, vw = g->last_vis_words;
    u64* rows = (u64*)(mm.arena + off_gather(g, g->last_parity));
    const bool fused = g->exchange != 1 || g->world == 1;    // peer stores (push kernel or epilogue)
    const bool epilogue = g->exchange == 2 && g->world > 1;
    // the kernel's visibility atomics go straight into this member's rows of its own gathered buffer (peer-store paths), or into
    // a dense local array that NCCL gathers afterwards
    u64* my_rows; size_t pitch_words;
    if (fused) { my_rows = rows + (size_t)mm.rank * rw; pitch_words = (size_t)g->world * rw; }
    else { my_rows = mm.d_dense.as<u64>(); pitch_words = rw; }
    if (n_my) {
        DMF_CUDA(cudaMemset2DAsync(my_rows, pitch_words * 8, 0, rw * 8, (size_t)n_my, st));
        DMF_CUDA(cudaMemsetAsync(mm.d_ticket.p, 0, ((size_t)n_my + 1) * 4, st));
    }
    PubTable pub; std::memset(&pub, 0, sizeof pub);
    if (fused) {
        pub.enabled = 1;
        pub.ticket = mm.d_ticket.as<unsigned>(); pub.views_done = mm.d_ticket.as<unsigned>() + n_my;
        pub.found_any = mm.d_found.as<int>();
        pub.n_views_pass = n_my; pub.row0 = mm.rank; pub.row_step = g->world;
        pub.row_words = (unsigned)rw; pub.vis_words = (unsigned)vw; pub.seq = g->seq;
        for (int r = 0; r < g->world; r++) {
            if (r == mm.rank) continue;
            pub.peer_rows[pub.n_peers] = (u64*)(mm.peer[r] + off_gather(g, g->last_parity));
            pub.peer_flag[pub.n_peers] = (unsigned*)(mm.peer[r] + kFlagGather + 4 * (size_t)mm.rank);
            pub.n_peers++;
        }
    }
    const bool march = n_my > 0 && (!spec.reverse || c->n_occ > 0);        // (the reverse march of an empty volume launches nothing)
    if (march) {
        DMF_TRY(order_after_previous(c, st));
        if (!spec.reverse) {
            FwdPlan pl; DMF_TRY(plan_forward(c, spec.fwd, pl));
            dmf_forward_out o{};
            if (extra) { o.depth_mm = extra->depth_mm; o.depth_u16 = extra->depth_u16; o.points = extra->points; o.hit_voxel = extra->hit_voxel; }
            o.visibility = (uint64_t*)my_rows; o.found_any = mm.d_found.as<int32_t>();
            c->defer_first_view = spec.fwd->mode == DMF_MODE_CLASSIFY && g->world > 1;    // resolved across GPUs by dmf_comm_fuse_marks
            const int rc = enqueue_forward(c, spec.fwd, pl, d_poses, n_my, spec.fwd->view_id0, o, nullptr, nullptr, nullptr, st, 0, nullptr, (unsigned)(pitch_words * 2), epilogue ? &pub : nullptr);
            c->defer_first_view = false;
            DMF_TRY(rc);
        } else {
            DMF_TRY(enqueue_reverse(c, spec.fast, spec.viz, d_poses, n_my, (unsigned*)my_rows, nullptr, mm.d_found.as<int>(), nullptr, nullptr, 0, st, (unsigned)(pitch_words * 2), epilogue ? &pub : nullptr));
        }
        if (fused && !epilogue) {            // the push kernel: found words into the rows, rows to the peers, flags
            k_publish_rows<<<n_my, 256, 0, st>>>(pub, my_rows, pitch_words);
            c->launches++;
            DMF_CUDA(cudaGetLastError());
        }
        DMF_TRY(mark_last(c, st));
    } else if (fused && g->world > 1) {
        DMF_TRY(comm_signal(g, mm, kFlagGather, g->seq, st));          // nothing to march: still tell the peers this rank is done
    }
    if (!fused && march) { k_found_into_rows<<<(n_my + 255) / 256, 256, 0, st>>>(my_rows, mm.d_found.as<int>(), n_my, (unsigned)rw, (unsigned)vw); c->launches++; DMF_CUDA(cudaGetLastError()); }
    (void)rows;
    return 0;
}

// NCCL path: all-gather of every member's dense rows (rank-major), then an interleave kernel into view order
int sweep_exchange_nccl(dmf_comm* g, int n_views, void* const* streams) {
    NcclApi* nc = nccl_api();
    const size_t rw = g->last_row_words;
    const int n_max = my_view_count(n_views, 0, g->world);
    if (g->single_process) DMF_NCCL(nc->GroupStart());
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = streams && streams[i] ? (cudaStream_t)streams[i] : mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        DMF_NCCL(nc->AllGather(mm.d_dense.p, mm.d_nccl_gather.p, (size_t)n_max * rw * 8, kNcclUint8, mm.nccl, st));
    }
    if (g->single_process) DMF_NCCL(nc->GroupEnd());
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = streams && streams[i] ? (cudaStream_t)streams[i] : mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        k_interleave_rows<<<blocks_for((size_t)n_views * rw, 256), 256, 0, st>>>(mm.d_nccl_gather.as<u64>(), (u64*)(mm.arena + off_gather(g, g->last_parity)), g->world, n_max, n_views, (unsigned)rw);
        mm.ctx->launches++;
        DMF_CUDA(cudaGetLastError());
    }
    return 0;
}

// Collective start of a sweep: sizes, sequence number, scratch of every local member.  Nothing is enqueued yet.
int sweep_begin(dmf_comm* g, int n_views) {
    if (n_views < 0) return fail("negative view count");
    DMF_TRY(comm_check_same_volume(g));
    const size_t vw = (g->m[0].ctx->n_occ + 63) / 64;
    const size_t rw = (vw + 2) & ~(size_t)1;                                      // + the found word, padded to 16 bytes
    // the staging region of the OR / min reduces is reserved along with the rows, so that fusing grids or marks after a sweep
    // never re-allocates the arena (which would drop the gathered rows)
    DMF_TRY(comm_reserve(g, std::max<size_t>((size_t)n_views * rw * 8, 256), reduce_stage_bytes(g)));
    g->seq++;
    g->last_parity = (int)(g->seq & 1u); g->last_n_views = n_views; g->last_row_words = rw; g->last_vis_words = vw; g->have_sweep = true;
    const int n_max = my_view_count(n_views, 0, g->world);
    for (auto& mm : g->m) {
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        DMF_TRY(mm.d_ticket.reserve(((size_t)n_max + 2) * 4)); DMF_TRY(mm.d_found.reserve(((size_t)n_max + 1) * 4)); DMF_TRY(mm.d_poses.reserve(((size_t)n_max + 1) * 48));
        if (g->exchange == 1 && g->world > 1) {
            DMF_TRY(mm.d_dense.reserve(std::max<size_t>((size_t)n_max * rw * 8, 256))); DMF_TRY(mm.d_nccl_gather.reserve(std::max<size_t>((size_t)g->world * n_max * rw * 8, 256)));
            DMF_CUDA(cudaMemsetAsync(mm.d_dense.p, 0, std::max<size_t>((size_t)n_max * rw * 8, 256), mm.ctx->stream));       // rows beyond n_my stay zero
        }
        // volume structures the march needs are built before anything is enqueued (they synchronise)
        if (mm.ctx->reverse_format == DMF_GRID_BYTE) DMF_TRY(ensure_bytes(mm.ctx, mm.ctx->stream));
    }
    return 0;
}

int sweep_run(dmf_comm* g, const SweepSpec& spec, const float* host_poses, const float* const* d_poses, int n_views, void* const* streams, const dmf_sweep_out* out) {
    DMF_TRY(sweep_begin(g, n_views));
    if (!spec.reverse) for (auto& mm : g->m) {               // tables / distance bytes synchronise: build them before the first enqueue
        FwdPlan pl; DMF_TRY(plan_forward(mm.ctx, spec.fwd, pl));
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        DMF_TRY(ensure_tables(mm.ctx, pl.z0, spec.fwd->zdelta, pl.cstride, pl.rstride, mm.ctx->stream));
        if (pl.grid_format == DMF_GRID_BYTE) DMF_TRY(ensure_bytes(mm.ctx, mm.ctx->stream));
        if (spec.fwd->flags & DMF_FWD_CARVE) DMF_TRY(ensure_observed(mm.ctx, mm.ctx->stream));
        DMF_TRY(mm.ctx->d_kstart.reserve((size_t)(my_view_count(n_views, 0, g->world) + 1) * 72 + 64));
    }
    else for (auto& mm : g->m) { DMF_CUDA(cudaSetDevice(mm.ctx->device)); DMF_TRY(mm.ctx->d_inv_poses.reserve((size_t)(my_view_count(n_views, 0, g->world) + 1) * 48)); }
    const size_t rw = g->last_row_words, vw = g->last_vis_words;
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = streams && streams[i] ? (cudaStream_t)streams[i] : mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        const int n_my = my_view_count(n_views, mm.rank, g->world);
        const float* dp = d_poses ? d_poses[i] : mm.d_poses.as<float>();
        if (!d_poses && n_my)            // this member's views out of the whole list: rows rank, rank + world, ...
            DMF_CUDA(cudaMemcpy2DAsync(mm.d_poses.p, 48, host_poses + 12 * (size_t)mm.rank, 48 * (size_t)g->world, 48, (size_t)n_my, cudaMemcpyHostToDevice, st));
        DMF_TRY(sweep_member_enqueue(g, mm, spec, dp, n_views, st, spec.d_out ? spec.d_out[i] : nullptr));
    }
    if (g->exchange == 1 && g->world > 1) DMF_TRY(sweep_exchange_nccl(g, n_views, streams));
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = streams && streams[i] ? (cudaStream_t)streams[i] : mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        const u64* rows = (const u64*)(mm.arena + off_gather(g, g->last_parity));
        // a single process sends every member's own rows over that member's own PCIe link; one process per GPU copies its own rows
        // or (DMF_SWEEP_ROWS_ALL) the whole gathered array
        const bool all = out && !g->single_process && out->rows_to_host == DMF_SWEEP_ROWS_ALL;
        auto rows_to_host = [&]() -> int {
            const int r0 = all ? 0 : mm.rank, step = all ? 1 : g->world, cnt = all ? n_views : mm.n_my;
            if (cnt > 0 && out->visibility && vw)
                DMF_CUDA(cudaMemcpy2DAsync(out->visibility + (size_t)r0 * vw, (size_t)step * vw * 8, rows + (size_t)r0 * rw, (size_t)step * rw * 8, vw * 8, (size_t)cnt, cudaMemcpyDeviceToHost, st));
            if (cnt > 0 && out->found_any)     // the low half of each row's extra word
                DMF_CUDA(cudaMemcpy2DAsync(out->found_any + r0, (size_t)step * 4, rows + (size_t)r0 * rw + vw, (size_t)step * rw * 8, 4, (size_t)cnt, cudaMemcpyDeviceToHost, st));
            return 0;
        };
        // A member's OWN rows are final as soon as its march (and push kernel) is: they go to the host before the stream waits for the
        // peers' rows, so the copy overlaps that wait -- and the ranks' copies, which share the host's memory path, spread out in time.
        // (The NCCL exchange interleaves the rows into view order after the all-gather: there the copy has to follow it.)
        const bool early = out && !all && g->exchange != 1;
        if (early) DMF_TRY(rows_to_host());
        if (g->exchange != 1) DMF_TRY(comm_wait(g, mm, kFlagGather, g->seq, st));
        if (out && !early) DMF_TRY(rows_to_host());
    }
    if (out) for (size_t i = 0; i < g->m.size(); i++) {
        cudaStream_t st = streams && streams[i] ? (cudaStream_t)streams[i] : g->m[i].ctx->stream;
        DMF_CUDA(cudaSetDevice(g->m[i].ctx->device));
        DMF_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---- OR / min reduce of a per-member buffer over the group ---------------------------------------------------------------
// bufs[i]: member i's device buffer of n32 32-bit words (same n32 everywhere).  OP 0: OR, 1: int32 min.  In place.
int comm_reduce(dmf_comm* g, const std::vector<unsigned*>& bufs, size_t n32, int op) {
    if (g->world == 1 || n32 == 0) return 0;
    const size_t n16 = (n32 + 3) / 4;                         // 16-byte units; the staging tail beyond n32 is padded with the identity
    DMF_TRY(comm_reserve(g, std::max<size_t>(g->gather_cap, 256), std::max(n16 * 16, reduce_stage_bytes(g))));
    const bool fused = g->exchange != 1;
    NcclApi* nc = nccl_api();
    if (!fused && op == 0) for (auto& mm : g->m) { DMF_CUDA(cudaSetDevice(mm.ctx->device)); DMF_TRY(mm.d_nccl_gather.reserve((size_t)g->world * n16 * 16)); }
    g->rseq++;
    if (!fused) {
        if (!nc) return fail("no peer access and no NCCL: cannot reduce over the group");
        if (g->single_process) DMF_NCCL(nc->GroupStart());
        for (size_t i = 0; i < g->m.size(); i++) {
            CommMember& mm = g->m[i];
            DMF_CUDA(cudaSetDevice(mm.ctx->device));
            if (op == 1) DMF_NCCL(nc->AllReduce(bufs[i], bufs[i], n32, kNcclInt32, kNcclMin, mm.nccl, mm.ctx->stream));
            else DMF_NCCL(nc->AllGather(bufs[i], mm.d_nccl_gather.p, n32 * 4, kNcclUint8, mm.nccl, mm.ctx->stream));      // NCCL has no bitwise OR
        }
        if (g->single_process) DMF_NCCL(nc->GroupEnd());
        if (op == 0) for (size_t i = 0; i < g->m.size(); i++) {
            CommMember& mm = g->m[i];
            DMF_CUDA(cudaSetDevice(mm.ctx->device));
            if (n32 % 2) return fail("OR reduce over NCCL needs an even word count");
            k_or_reduce<<<blocks_for(n32 / 2, 256, 148 * 8), 256, 0, mm.ctx->stream>>>((u64*)bufs[i], mm.d_nccl_gather.as<u64>(), g->world, n32 / 2);
            mm.ctx->launches++;
        }
        for (auto& mm : g->m) { DMF_CUDA(cudaSetDevice(mm.ctx->device)); DMF_CUDA(cudaStreamSynchronize(mm.ctx->stream)); }
        return 0;
    }
    // 1. stage, 2. tell the peers
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        char* stage = mm.arena + off_stage(g);
        if (n16 * 4 > n32) DMF_TRY(fill_u32(mm.ctx, st, stage + (n16 - 1) * 16, 4, op == 0 ? 0u : 0x7fffffffu));
        DMF_CUDA(cudaMemcpyAsync(stage, bufs[i], n32 * 4, cudaMemcpyDeviceToDevice, st));
        DMF_TRY(comm_signal(g, mm, kFlagStaged, g->rseq, st));
    }
    // 3. wait for everybody's staging, 4. reduce this rank's slice from all peers and push it back, 5. tell the peers
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        DMF_TRY(comm_wait(g, mm, kFlagStaged, g->rseq, st));
        PeerReduceArgs a; a.world = g->world;
        a.stage[0] = (unsigned*)(mm.arena + off_stage(g));                  // own copy first, then the peers
        int k = 1;
        for (int r = 0; r < g->world; r++) if (r != mm.rank) a.stage[k++] = (unsigned*)(mm.peer[r] + off_stage(g));
        a.lo = n16 * (size_t)mm.rank / g->world; a.hi = n16 * (size_t)(mm.rank + 1) / g->world;
        if (a.hi > a.lo) {
            if (op == 0) k_peer_reduce<0><<<blocks_for(a.hi - a.lo, 256, 148 * 8), 256, 0, st>>>(a);
            else k_peer_reduce<1><<<blocks_for(a.hi - a.lo, 256, 148 * 8), 256, 0, st>>>(a);
            mm.ctx->launches++;
            DMF_CUDA(cudaGetLastError());
        }
        DMF_TRY(comm_signal(g, mm, kFlagReduced, g->rseq, st));
    }
    // 6. wait for every slice, 7. take the result
    for (size_t i = 0; i < g->m.size(); i++) {
        CommMember& mm = g->m[i];
        cudaStream_t st = mm.ctx->stream;
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        DMF_TRY(comm_wait(g, mm, kFlagReduced, g->rseq, st));
        DMF_CUDA(cudaMemcpyAsync(bufs[i], mm.arena + off_stage(g), n32 * 4, cudaMemcpyDeviceToDevice, st));
    }
    for (auto& mm : g->m) { DMF_CUDA(cudaSetDevice(mm.ctx->device)); DMF_CUDA(cudaStreamSynchronize(mm.ctx->stream)); }
    return 0;
}

int comm_finish_init(dmf_comm* g) {
    // cuStreamWaitValue32 through the runtime's driver entry point query: no link-time dependency on libcuda
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (!std::getenv("DMF_COMM_SPIN_WAIT") && cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess && fn)
        g->wait32 = (WaitValue32Fn)fn;
    else cudaGetLastError();
    const char* ex = std::getenv("DMF_COMM_EXCHANGE");
    if (ex && std::string(ex) == "nccl") g->p2p = false;             // forced: the A/B baseline (must be set on every rank alike)
    g->exchange = g->p2p ? ((ex && std::string(ex) == "epilogue") ? 2 : 0) : 1;
    if (g->world > DMF_MAX_PEERS) return fail("at most %d GPUs per group (got %d)", DMF_MAX_PEERS, g->world);
    if (g->exchange == 1 && g->world > 1 && (!nccl_api() || !g->m[0].nccl)) return fail("GPUs cannot address each other and NCCL is not available (%s)", nccl_api() ? "no communicator" : "dlopen failed");
    for (auto& mm : g->m) mm.peer.assign(g->world, nullptr);
    return 0;
}

}  // namespace

extern "C" {

int dmf_comm_unique_id(void* id) {
    if (!id) return fail("null argument");
    NcclApi* nc = nccl_api();
    if (!nc) return fail("NCCL not available: dlopen(libnccl.so.2) failed");
    NcclId nid;
    DMF_NCCL(nc->GetUniqueId(&nid));
    std::memcpy(id, &nid, sizeof nid);
    return 0;
}

int dmf_comm_init_all(dmf_comm** out, int n_gpus) {
    if (!out) return fail("null argument");
    *out = nullptr;
    const int have = dmf_device_count();
    if (have <= 0) return fail("no CUDA device visible: libdmf_b200 has no CPU fallback");
    if (n_gpus <= 0) n_gpus = std::min(have, DMF_MAX_PEERS);
    if (n_gpus > have) return fail("%d GPUs requested, %d visible", n_gpus, have);
    if (n_gpus > DMF_MAX_PEERS) return fail("at most %d GPUs per group", DMF_MAX_PEERS);
    dmf_comm* g = new (std::nothrow) dmf_comm();
    if (!g) return fail("out of host memory");
    g->world = g->n_local = n_gpus; g->rank0 = 0; g->single_process = true; g->owns_ctx = true;
    g->m.resize(n_gpus);
    for (int i = 0; i < n_gpus; i++) {
        g->m[i].rank = i;
        if (dmf_create(&g->m[i].ctx, i)) { dmf_comm_destroy(g); return 1; }
    }
    // peer access between every pair
    g->p2p = true;
    for (int i = 0; i < n_gpus && g->p2p; i++) for (int j = 0; j < n_gpus; j++) {
        if (i == j) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, i, j) != cudaSuccess || !can) { cudaGetLastError(); g->p2p = false; break; }
    }
    if (g->p2p) for (int i = 0; i < n_gpus; i++) {
        cudaSetDevice(i);
        for (int j = 0; j < n_gpus; j++) {
            if (i == j) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); g->p2p = false; }
        }
    }
    // one process, peers addressable: nothing needs NCCL (and ncclCommInitAll costs seconds); it is formed only as the fallback
    // exchange, or when DMF_COMM_EXCHANGE=nccl asks for the A/B baseline
    const char* ex_env = std::getenv("DMF_COMM_EXCHANGE");
    const bool want_nccl = !g->p2p || (ex_env && std::string(ex_env) == "nccl");
    if (n_gpus > 1 && want_nccl) {
        if (NcclApi* nc = nccl_api()) {
            std::vector<NcclComm> comms(n_gpus, nullptr); std::vector<int> devs(n_gpus);
            for (int i = 0; i < n_gpus; i++) devs[i] = i;
            const int r = nc->CommInitAll(comms.data(), n_gpus, devs.data());
            if (r == 0) for (int i = 0; i < n_gpus; i++) g->m[i].nccl = comms[i];
            else if (!g->p2p) { fail("ncclCommInitAll failed: %s", nc->GetErrorString(r)); dmf_comm_destroy(g); return 1; }
        }
    }
    if (comm_finish_init(g)) { dmf_comm_destroy(g); return 1; }
    *out = g;
    return 0;
}

int dmf_comm_init_rank(dmf_comm** out, dmf_ctx* ctx, const void* unique_id, int rank, int world) {
    if (!out || !ctx) return fail("null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank %d / world %d", rank, world);
    if (world > DMF_MAX_PEERS) return fail("at most %d GPUs per group", DMF_MAX_PEERS);
    dmf_comm* g = new (std::nothrow) dmf_comm();
    if (!g) return fail("out of host memory");
    g->world = world; g->n_local = 1; g->rank0 = rank; g->single_process = false; g->owns_ctx = false;
    g->m.resize(1);
    g->m[0].ctx = ctx; g->m[0].rank = rank;
    g->p2p = world > 1;                       // CUDA IPC between the GPUs of one box; verified when the arenas are mapped
    if (world > 1) {
        NcclApi* nc = nccl_api();
        if (!nc) { delete g; return fail("NCCL not available: dlopen(libnccl.so.2) failed"); }
        if (!unique_id) { delete g; return fail("null unique id"); }
        NcclId nid; std::memcpy(&nid, unique_id, sizeof nid);
        cudaSetDevice(ctx->device);
        const int r = nc->CommInitRank(&g->m[0].nccl, world, nid, rank);
        if (r != 0) { delete g; return fail("ncclCommInitRank failed: %s", nc->GetErrorString(r)); }
    }
    if (comm_finish_init(g)) { dmf_comm_destroy(g); return 1; }
    *out = g;
    return 0;
}

void dmf_comm_destroy(dmf_comm* g) {
    if (!g) return;
    if (!g->m.empty() && g->m[0].ctx) comm_free_arenas(g);
    NcclApi* nc = nccl_api();
    for (auto& mm : g->m) {
        if (!mm.ctx) continue;
        cudaSetDevice(mm.ctx->device);
        if (mm.nccl && nc) nc->CommDestroy(mm.nccl);
        for (DevBuf* b : {&mm.d_ticket, &mm.d_poses, &mm.d_found, &mm.d_dense, &mm.d_nccl_gather, &mm.d_hdr, &mm.d_ranks}) b->release();
        if (g->owns_ctx) dmf_destroy(mm.ctx);
    }
    delete g;
}

int dmf_comm_info(dmf_comm* g, int* world, int* n_local, int* first_rank, int* exchange) {
    if (!g) return fail("null group");
    if (world) *world = g->world;
    if (n_local) *n_local = g->n_local;
    if (first_rank) *first_rank = g->rank0;
    if (exchange) *exchange = g->world == 1 ? DMF_EXCHANGE_NONE : (g->exchange == 0 ? DMF_EXCHANGE_P2P_PUSH : (g->exchange == 1 ? DMF_EXCHANGE_NCCL : DMF_EXCHANGE_P2P_EPILOGUE));
    return 0;
}

dmf_ctx* dmf_comm_ctx(dmf_comm* g, int local_index) {
    if (!g || local_index < 0 || local_index >= (int)g->m.size()) return nullptr;
    return g->m[local_index].ctx;
}

int dmf_comm_set_camera(dmf_comm* g, const float K[9], int height, int width) {
    if (!g) return fail("null group");
    for (auto& mm : g->m) DMF_TRY(dmf_set_camera(mm.ctx, K, height, width));
    return 0;
}

// The volume uploaded on rank `root` goes to every member GPU to GPU: the occupied id list and the normals CSR travel
// (cudaMemcpyPeer within a process, ncclBroadcast between processes), each GPU rebuilds the march structures locally.
int dmf_comm_replicate_volume(dmf_comm* g, int root) {
    if (!g) return fail("null group");
    if (root < 0 || root >= g->world) return fail("bad root %d", root);
    if (g->world == 1) return g->m[0].ctx->vol_set ? 0 : fail("no volume uploaded");
    struct Hdr { double bounds[6], delta[3]; int dim[3]; int pad; unsigned long long n_occ, n_normals; } h;
    std::memset(&h, 0, sizeof h);
    CommMember* rm = nullptr;
    for (auto& mm : g->m) if (mm.rank == root) rm = &mm;
    auto fill_hdr = [&](dmf_ctx* c) { std::memcpy(h.bounds, c->bounds, sizeof h.bounds); for (int a = 0; a < 3; a++) { h.delta[a] = c->vol.delta[a]; h.dim[a] = c->vol.dim[a]; } h.n_occ = c->n_occ; h.n_normals = c->n_normals; };
    if (g->single_process) {
        if (!rm->ctx->vol_set) return fail("the root has no volume uploaded");
        fill_hdr(rm->ctx);
        DMF_CUDA(cudaSetDevice(rm->ctx->device)); DMF_CUDA(cudaStreamSynchronize(rm->ctx->stream));
        for (auto& mm : g->m) {
            if (&mm == rm) continue;
            dmf_ctx* c = mm.ctx;
            DMF_CUDA(cudaSetDevice(c->device)); DMF_CUDA(cudaDeviceSynchronize());
            DMF_TRY(set_volume_geometry(c, h.bounds, h.delta, h.dim));
            c->vol_set = false; c->mirror_valid = false;
            DMF_TRY(c->d_occ_ids.reserve(std::max<size_t>(h.n_occ, 1) * 8)); DMF_TRY(c->d_noff.reserve((h.n_occ + 1) * 4)); DMF_TRY(c->d_normals.reserve(std::max<size_t>(3 * h.n_normals, 1) * 4));
            if (h.n_occ) DMF_CUDA(cudaMemcpyPeerAsync(c->d_occ_ids.p, c->device, rm->ctx->d_occ_ids.p, rm->ctx->device, h.n_occ * 8, c->stream));
            DMF_CUDA(cudaMemcpyPeerAsync(c->d_noff.p, c->device, rm->ctx->d_noff.p, rm->ctx->device, (h.n_occ + 1) * 4, c->stream));
            if (h.n_normals) DMF_CUDA(cudaMemcpyPeerAsync(c->d_normals.p, c->device, rm->ctx->d_normals.p, rm->ctx->device, 3 * h.n_normals * 4, c->stream));
            DMF_TRY(build_volume_device(c, h.n_occ, h.n_normals));
        }
        return 0;
    }
    CommMember& mm = g->m[0];
    dmf_ctx* c = mm.ctx;
    NcclApi* nc = nccl_api();
    if (!nc || !mm.nccl) return fail("multi-process group without NCCL");
    DMF_CUDA(cudaSetDevice(c->device));
    const bool is_root = mm.rank == root;
    if (is_root) { if (!c->vol_set) return fail("the root has no volume uploaded"); fill_hdr(c); }
    DMF_TRY(mm.d_hdr.reserve(std::max<size_t>(sizeof h, 4096)));
    if (is_root) DMF_CUDA(cudaMemcpyAsync(mm.d_hdr.p, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    DMF_NCCL(nc->Broadcast(mm.d_hdr.p, mm.d_hdr.p, sizeof h, kNcclUint8, root, mm.nccl, c->stream));
    DMF_CUDA(cudaMemcpyAsync(&h, mm.d_hdr.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    if (!is_root) {
        DMF_CUDA(cudaDeviceSynchronize());
        DMF_TRY(set_volume_geometry(c, h.bounds, h.delta, h.dim));
        c->vol_set = false; c->mirror_valid = false;
        DMF_TRY(c->d_occ_ids.reserve(std::max<size_t>(h.n_occ, 1) * 8)); DMF_TRY(c->d_noff.reserve((h.n_occ + 1) * 4)); DMF_TRY(c->d_normals.reserve(std::max<size_t>(3 * h.n_normals, 1) * 4));
    }
    if (h.n_occ) DMF_NCCL(nc->Broadcast(c->d_occ_ids.p, c->d_occ_ids.p, h.n_occ * 8, kNcclUint8, root, mm.nccl, c->stream));
    DMF_NCCL(nc->Broadcast(c->d_noff.p, c->d_noff.p, (h.n_occ + 1) * 4, kNcclUint8, root, mm.nccl, c->stream));
    if (h.n_normals) DMF_NCCL(nc->Broadcast(c->d_normals.p, c->d_normals.p, 3 * h.n_normals * 4, kNcclUint8, root, mm.nccl, c->stream));
    DMF_CUDA(cudaStreamSynchronize(c->stream));
    if (!is_root) DMF_TRY(build_volume_device(c, h.n_occ, h.n_normals));
    return 0;
}

int dmf_sweep_forward(dmf_comm* g, const dmf_forward_params* p, const float* poses, int n_views, const dmf_sweep_out* out) {
    if (!g || !p || (!poses && n_views > 0)) return fail("null argument");
    if (p->mode == DMF_MODE_MINIMUM || p->mode == DMF_MODE_MARK) return fail("dmf_sweep_forward supports POINTS, GOOD_POINTS and CLASSIFY");
    SweepSpec s; s.reverse = false; s.fwd = p;
    return sweep_run(g, s, poses, nullptr, n_views, nullptr, out);
}
int dmf_sweep_forward_dev(dmf_comm* g, const dmf_forward_params* p, const float* const* d_poses, int n_views, const dmf_forward_out* const* d_out, void* const* streams) {
    if (!g || !p || !d_poses) return fail("null argument");
    if (p->mode == DMF_MODE_MINIMUM || p->mode == DMF_MODE_MARK) return fail("dmf_sweep_forward supports POINTS, GOOD_POINTS and CLASSIFY");
    SweepSpec s; s.reverse = false; s.fwd = p; s.d_out = d_out;
    return sweep_run(g, s, nullptr, d_poses, n_views, streams, nullptr);
}
int dmf_sweep_reverse(dmf_comm* g, int fast, const float* poses, int n_views, const dmf_sweep_out* out) {
    if (!g || (!poses && n_views > 0)) return fail("null argument");
    if (!fast) return fail("dmf_sweep_reverse: only reverseRayTraceFast (fast = 1) is sharded; use dmf_reverse for the whole-grid scan");
    SweepSpec s; s.reverse = true; s.fast = 1; s.viz = 0;
    return sweep_run(g, s, poses, nullptr, n_views, nullptr, out);
}
int dmf_sweep_reverse_dev(dmf_comm* g, int fast, const float* const* d_poses, int n_views, void* const* streams) {
    if (!g || !d_poses) return fail("null argument");
    if (!fast) return fail("dmf_sweep_reverse: only reverseRayTraceFast (fast = 1) is sharded");
    SweepSpec s; s.reverse = true; s.fast = 1; s.viz = 0;
    return sweep_run(g, s, nullptr, d_poses, n_views, streams, nullptr);
}

int dmf_sweep_gathered_dev(dmf_comm* g, int local_index, uint64_t** d_rows, size_t* row_words, size_t* vis_words, int* n_views) {
    if (!g || local_index < 0 || local_index >= (int)g->m.size()) return fail("bad group / member");
    if (!g->have_sweep) return fail("no sweep yet");
    if (d_rows) *d_rows = (uint64_t*)(g->m[local_index].arena + off_gather(g, g->last_parity));
    if (row_words) *row_words = g->last_row_words;
    if (vis_words) *vis_words = g->last_vis_words;
    if (n_views) *n_views = g->last_n_views;
    return 0;
}

int dmf_sweep_set_cover(dmf_comm* g, int32_t* selected, int* n_selected) {
    if (!g || !selected || !n_selected) return fail("null argument");
    if (!g->have_sweep) return fail("no sweep yet");
    CommMember& mm = g->m[0];
    DMF_CUDA(cudaSetDevice(mm.ctx->device));
    DMF_CUDA(cudaStreamSynchronize(mm.ctx->stream));
    return greedy_set_cover_strided(mm.ctx, (const uint64_t*)(mm.arena + off_gather(g, g->last_parity)), g->last_n_views, g->last_vis_words, g->last_row_words, selected, n_selected);
}

// carve mode over the group: every member's observed grid becomes the union
int dmf_comm_fuse_observed(dmf_comm* g) {
    if (!g) return fail("null group");
    DMF_TRY(comm_check_same_volume(g));
    std::vector<unsigned*> bufs;
    for (auto& mm : g->m) {
        DMF_CUDA(cudaSetDevice(mm.ctx->device));
        DMF_TRY(ensure_observed(mm.ctx, mm.ctx->stream));
        DMF_CUDA(cudaDeviceSynchronize());
        bufs.push_back(mm.ctx->d_observed.as<unsigned>());
    }
    return comm_reduce(g, bufs, g->m[0].ctx->n_grid_words, 0);
}

// After a sharded CLASSIFY sweep (dmf_sweep_forward, mode CLASSIFY, view_id0 = id of view 0): Voxel::view = view_id0 + the
// smallest global view index that hit the voxel, where it was still 0 (first-wins in call order, RayTracingEngine.hpp:354-355);
// Voxel::good = OR over the GPUs (:356-370).  Afterwards every member holds the same marks.
int dmf_comm_fuse_marks(dmf_comm* g, int view_id0) {
    if (!g) return fail("null group");
    DMF_TRY(comm_check_same_volume(g));
    const size_t n_occ = g->m[0].ctx->n_occ;
    if (!n_occ) return 0;
    std::vector<unsigned*> fv, gb;
    for (auto& mm : g->m) {
        dmf_ctx* c = mm.ctx;
        DMF_CUDA(cudaSetDevice(c->device));
        DMF_CUDA(cudaDeviceSynchronize());
        if (g->world > 1) { k_first_view_to_global<<<blocks_for(n_occ, 256, 1u << 30), 256, 0, c->stream>>>(c->d_first_view.as<int>(), (int)n_occ, mm.rank, g->world); c->launches++; }
        fv.push_back(c->d_first_view.as<unsigned>()); gb.push_back(c->d_good_bits.as<unsigned>());
    }
    DMF_TRY(comm_reduce(g, fv, n_occ, 1));
    DMF_TRY(comm_reduce(g, gb, ((n_occ + 63) / 64) * 2, 0));
    for (auto& mm : g->m) {
        dmf_ctx* c = mm.ctx;
        DMF_CUDA(cudaSetDevice(c->device));
        k_apply_first_view<<<blocks_for(n_occ, 256, 1u << 30), 256, 0, c->stream>>>(c->d_view_mark.as<int>(), c->d_first_view.as<int>(), (int)n_occ, view_id0);
        c->launches++;
        DMF_CUDA(cudaGetLastError());
        DMF_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int dmf_comm_synchronize(dmf_comm* g) {
    if (!g) return fail("null group");
    for (auto& mm : g->m) { DMF_CUDA(cudaSetDevice(mm.ctx->device)); DMF_CUDA(cudaDeviceSynchronize()); }
    return 0;
}

}  // extern "C"
