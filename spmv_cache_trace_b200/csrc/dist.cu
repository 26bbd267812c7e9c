// dist.cu -- the row-partitioned multi-GPU mode behind the C ABI (spmvb200_comm_*, spmvb200_dist_*).
//
// The reference is one process whose OpenMP threads own ceil(rows/T) consecutive rows each
// (matrix/csr-matrix.cpp:77-95) and meet at barriers around Kernel::run (profile-kernel.cpp:159-161); its model of
// "who owns x_j" is the first-touch page owner (util/aligned-allocator.hpp:156-211).  Here a rank owns a row block on
// its own GPU and the matching slice of x, and the iteration x_(k+1) = alpha*A*x_k needs one exchange of x per step.
//
// Per rank and step k (buffers: cur = X[k%2] holds x_k, nxt = X[(k+1)%2] receives the rank's rows of x_(k+1)):
//
//   stream "comm" (high priority)  wait: the rank's slice of x_k is complete (E_int[k-1], E_bnd[k-1])
//                                  exchange: fill the remote columns of cur this rank's rows reference
//                                  record E_exch[k]
//   stream "int"                   wait: E_bnd[k-1] (stream order gives E_int[k-1]); nobody still reads nxt's slice
//                                  blocks whose rows reference only the rank's own columns  -> nxt
//                                  record E_int[k]
//   stream "bnd"  (high priority)  wait: E_exch[k], E_int[k-1]; (pieces that ADD to rows another piece stores: E_int[k])
//                                  blocks that reference remote columns                     -> nxt
//                                  record E_bnd[k]
//
// so the boundary rows run concurrently with the interior block as soon as their halo has arrived, the next
// exchange starts as soon as the boundary rows are done, and a step costs max(interior, exchange + boundary).
// Every block computes y = alpha*A*x with plain stores where a thread owns whole rows ("beta0", spmvb200_set_alpha).
//
// Exchange backends:
//   in-process  ranks share the address space: rank p PULLS the ranges it needs from the owners' buffers with
//               cudaMemcpyPeerAsync (NVLink P2P between devices, a plain copy on one device), after waiting on the
//               owner's events; the owner's next-but-one step waits on the puller's E_exch before it overwrites
//               the slice (write-after-read across ranks).
//   NCCL        one process per rank: ncclSend/ncclRecv of exactly the needed ranges in one group ("halo"), or
//               ncclAllGather / grouped ncclBroadcast of the slices ("all-gather").  libnccl.so.2 is dlopen'ed.
//   peer copy   one process per rank, flag SPMVB200_DIST_PEER_COPY: the in-process scheme across processes.  Every rank
//               exports its x buffers (cudaIpcGetMemHandle) and its "x_k is ready" / "my pulls of step k are done" events
//               (interprocess events); a rank pulls its ranges straight out of the owners' buffers with cudaMemcpyAsync on
//               the peer-mapped pointers -- the copy engines move the data over NVLink, no SM is taken from the SpMV, no
//               rendezvous with the sender.  An interprocess event can only be waited on once its record has been ISSUED,
//               so the ranks publish their host progress (steps whose events are recorded) in a small POSIX shared-memory
//               block and a waiter spins on the HOST until the owner's counter says the record exists; nothing spins on
//               the device.  NCCL is used only while the executor is set up (handle exchange, barriers).
#include "common.cuh"

#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#define SPMV_ABI_CATCH                                                                                  \
    catch (const std::bad_alloc &) { return ::spmvb200::fail(SPMVB200_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception & e) { return ::spmvb200::fail(SPMVB200_ERR_INVALID, e.what()); }

namespace spmvb200 {

// ---- NCCL, loaded on first use ------------------------------------------------------------------------------------
struct NcclApi {
    void * handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char * (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi * nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a process that already holds an NCCL (e.g. torch's bundled one) gets that one: same SONAME
        for (const char * name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        bool ok = true;
        auto sym = [&](const char * n) {
            void * p = dlsym(api.handle, n);
            if (!p) { ok = false; api.error = std::string("libnccl lacks ") + n; }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return &api;
}

static int nccl_fail(ncclResult_t r, const char * what)
{
    NcclApi * n = nccl_api();
    return fail(SPMVB200_ERR_CUDA, std::string("NCCL error: ") + (n->GetErrorString ? n->GetErrorString(r) : "?") + " in " + what);
}
#define SPMV_NCCL(call)                                                  \
    do {                                                                 \
        ncclResult_t r__ = (call);                                       \
        if (r__ != ncclSuccess) return ::spmvb200::nccl_fail(r__, #call); \
    } while (0)

struct Range {
    int peer;
    int64_t lo, hi;
};

// ---- the exchange plan: plain arithmetic --------------------------------------------------------------------------
struct Plan {
    int mode = SPMVB200_EXCHANGE_ALLGATHER;
    std::vector<Range> sends, recvs;
    int64_t recv_bytes = 0, send_bytes = 0;
};

// What rank q must receive: need[q] minus its own slice, cut at the owners' boundaries.
static void remote_ranges(int parts, const int64_t * starts, int64_t lo, int64_t hi, int q, std::vector<Range> & out)
{
    if (hi <= lo) return;
    for (int o = 0; o < parts; o++) {
        if (o == q) continue;
        const int64_t a = std::max(lo, starts[o]), b = std::min(hi, starts[o + 1]);
        if (b > a) out.push_back({o, a, b});
    }
}

static Plan make_plan(int parts, const int64_t * starts, const int64_t * need_lo, const int64_t * need_hi, int rank, int mode)
{
    Plan plan;
    const int64_t n = starts[parts];
    if (mode == SPMVB200_EXCHANGE_AUTO) {
        int64_t worst = 0;
        for (int q = 0; q < parts; q++) {
            std::vector<Range> r;
            remote_ranges(parts, starts, need_lo[q], need_hi[q], q, r);
            int64_t tot = 0;
            for (auto & t : r) tot += t.hi - t.lo;
            worst = std::max(worst, tot);
        }
        mode = 4 * worst <= n ? SPMVB200_EXCHANGE_HALO : SPMVB200_EXCHANGE_ALLGATHER;
    }
    plan.mode = mode;
    if (mode == SPMVB200_EXCHANGE_ALLGATHER) {
        for (int o = 0; o < parts; o++) {
            if (o == rank) continue;
            if (starts[o + 1] > starts[o]) plan.recvs.push_back({o, starts[o], starts[o + 1]});
            if (starts[rank + 1] > starts[rank]) plan.sends.push_back({o, starts[rank], starts[rank + 1]});
        }
    } else {
        remote_ranges(parts, starts, need_lo[rank], need_hi[rank], rank, plan.recvs);
        for (int q = 0; q < parts; q++) {
            if (q == rank) continue;
            std::vector<Range> r;
            remote_ranges(parts, starts, need_lo[q], need_hi[q], q, r);
            for (auto & t : r)
                if (t.peer == rank) plan.sends.push_back({q, t.lo, t.hi});
        }
    }
    for (auto & t : plan.recvs) plan.recv_bytes += 8 * (t.hi - t.lo);
    for (auto & t : plan.sends) plan.send_bytes += 8 * (t.hi - t.lo);
    return plan;
}

}  // namespace spmvb200

using namespace spmvb200;

// ---- communicators ----------------------------------------------------------------------------------------------------

struct LocalGroup {
    std::mutex mu;
    std::condition_variable cv;
    int nranks = 0;
    std::vector<spmvb200_dist_s *> dist;   // registered executors, by rank
    std::vector<int64_t> need_lo, need_hi;  // column range each rank's rows reference
    std::vector<int> need_set;
    int wanted_mode = -1;
    // thread rendezvous (barrier / allreduce, one thread per rank)
    std::vector<double> vals;
    int arrived = 0;
    uint64_t generation = 0;
    double result = 0.0;
};

struct spmvb200_comm_s {
    int rank = 0, nranks = 1, device = 0;
    bool local = true;
    std::shared_ptr<LocalGroup> group;
    ncclComm_t nccl = nullptr;
    cudaStream_t stream = nullptr;
    double * scratch = nullptr;  // a few doubles of device memory for the plumbing collectives
    uint64_t id_hash = 0;        // names the shared-memory block of the peer-copy transport
    int executors = 0;           // executors created on this communicator so far (part of that name)
};

struct DistBlock {
    spmvb200_matrix_t A = nullptr;
    int64_t b = 0, e = 0;
    bool remote = false, accumulate = false, owned = true;
};

constexpr int kRing = 4;  // event slots: step k uses slot k % kRing
constexpr int kPullStreams = 4;  // peer-copy transport: pulls in flight at once (one copy engine each)

struct spmvb200_dist_s {
    spmvb200_comm_t comm = nullptr;
    int rank = 0, P = 1, device = 0;
    std::vector<int64_t> starts;
    int64_t n = 0, s = 0, e = 0, rows = 0, nnz = 0;
    double * X[2] = {nullptr, nullptr};
    double * Yh[2] = {nullptr, nullptr};  // run_host: the rank's rows of y, double-buffered
    cudaStream_t s_int = nullptr, s_bnd = nullptr, s_comm = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t e_exch[kRing] = {}, e_int[kRing] = {}, e_bnd[kRing] = {}, e_up[kRing] = {}, e_down[kRing] = {};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::vector<DistBlock> blocks;
    Plan plan;
    bool plan_ready = false;
    int wanted_mode = 0;
    bool any_accumulate = false;
    int64_t need_lo = 0, need_hi = 0;
    int64_t k = 0;  // steps issued
    int64_t x_bytes = 0;
    bool equal_slices = false;
    // peer-copy transport (one process per rank)
    bool peer = false;
    cudaEvent_t e_xready[kRing] = {};        // interprocess: "this rank's slice of x_k is complete", slot k % kRing
    std::vector<double *> peer_X[2];         // [buffer][rank] peer-mapped x buffers (own rank: the local pointer)
    std::vector<cudaEvent_t> peer_xready[kRing], peer_exch[kRing];  // [slot][rank] opened interprocess events
    cudaStream_t s_pull[kPullStreams] = {};  // extra streams of the all-gather pulls
    int pull_lanes = 1;                      // how many of them are used (SPMVB200_PULL_LANES, 1..kPullStreams)
    cudaEvent_t e_fork = nullptr, e_join[kPullStreams] = {};
    // fused halo push: the kernels that compute the rows a neighbour references store them into the neighbour's buffer
    struct PushRange { int peer; int64_t lo, hi; };          // block-local rows [lo, hi) go to rank `peer`
    std::vector<std::vector<PushRange>> push;                // per block
    bool want_push = false, push_active = false;
    int64_t halo_by_push = -1;               // the halo of x_k (k = this value) was delivered by the senders' kernels
    struct PeerCounters * shm = nullptr;     // [P] host progress of every rank, in POSIX shared memory
    size_t shm_bytes = 0;
};

// Host progress a rank publishes for the others: events up to these step indices have been RECORDED (issued).
struct PeerCounters {
    volatile int64_t xready_issued;  // e_xready[k] exists for all k < xready_issued
    volatile int64_t exch_issued;    // e_exch[k]   exists for all k < exch_issued
    char pad[48];
};

namespace {

int comm_check(spmvb200_comm_t c)
{
    if (!c) return fail(SPMVB200_ERR_INVALID, "null communicator");
    SPMV_CUDA(cudaSetDevice(c->device));
    return 0;
}

int dist_check(spmvb200_dist_t d)
{
    if (!d) return fail(SPMVB200_ERR_INVALID, "null executor handle");
    SPMV_CUDA(cudaSetDevice(d->device));
    return 0;
}

// one thread per rank meets here; returns the reduction of `v` over the ranks
double group_reduce(LocalGroup & g, int rank, double v, int op)
{
    std::unique_lock<std::mutex> lk(g.mu);
    const uint64_t gen = g.generation;
    g.vals[(size_t)rank] = v;
    if (++g.arrived == g.nranks) {
        double r = g.vals[0];
        for (int q = 1; q < g.nranks; q++) {
            const double t = g.vals[(size_t)q];
            r = op == 0 ? std::max(r, t) : op == 1 ? r + t : std::min(r, t);
        }
        g.result = r;
        g.arrived = 0;
        g.generation++;
        g.cv.notify_all();
        return r;
    }
    // (a rank that failed and never arrives must not hang the others for ever)
    g.cv.wait_for(lk, std::chrono::seconds(300), [&] { return g.generation != gen; });
    return g.result;
}

void finalize_plan(spmvb200_dist_t d, const int64_t * need_lo, const int64_t * need_hi)
{
    d->plan = make_plan(d->P, d->starts.data(), need_lo, need_hi, d->rank, d->wanted_mode);
    d->plan_ready = true;
}

int dist_free(spmvb200_dist_t d)
{
    if (!d) return 0;
    cudaSetDevice(d->device);
    for (cudaStream_t s : {d->s_int, d->s_bnd, d->s_comm, d->s_h2d, d->s_d2h})
        if (s) cudaStreamSynchronize(s);
    if (d->comm && d->comm->local && d->comm->group) {
        // the other ranks of the process pull from (and, with the fused push, store into) this rank's buffers: whatever
        // they have queued must be done before the buffers go away
        std::vector<spmvb200_dist_t> others;
        {
            std::lock_guard<std::mutex> lk(d->comm->group->mu);
            for (spmvb200_dist_t o : d->comm->group->dist)
                if (o && o != d) others.push_back(o);
            if ((size_t)d->rank < d->comm->group->dist.size() && d->comm->group->dist[(size_t)d->rank] == d) {
                d->comm->group->dist[(size_t)d->rank] = nullptr;
                d->comm->group->need_set[(size_t)d->rank] = 0;
            }
            if (others.empty()) d->comm->group->wanted_mode = -1;  // the last executor of this generation
        }
        for (spmvb200_dist_t o : others) {
            cudaSetDevice(o->device);
            for (cudaStream_t s : {o->s_comm, o->s_int, o->s_bnd})
                if (s) cudaStreamSynchronize(s);
        }
        cudaSetDevice(d->device);
    }
    if (d->peer) {
        for (int q = 0; q < d->P; q++) {
            if (q == d->rank) continue;
            for (int b = 0; b < 2; b++)
                if ((size_t)q < d->peer_X[b].size() && d->peer_X[b][(size_t)q]) cudaIpcCloseMemHandle(d->peer_X[b][(size_t)q]);
            for (int i = 0; i < kRing; i++) {
                if ((size_t)q < d->peer_xready[i].size() && d->peer_xready[i][(size_t)q]) cudaEventDestroy(d->peer_xready[i][(size_t)q]);
                if ((size_t)q < d->peer_exch[i].size() && d->peer_exch[i][(size_t)q]) cudaEventDestroy(d->peer_exch[i][(size_t)q]);
            }
        }
        if (d->comm && !d->comm->local && d->P > 1) spmvb200_comm_barrier(d->comm);  // nobody frees a buffer a peer still maps
    }
    if (d->shm) munmap((void *)d->shm, d->shm_bytes);
    for (int l = 0; l < kPullStreams; l++) {
        if (d->s_pull[l]) { cudaStreamSynchronize(d->s_pull[l]); cudaStreamDestroy(d->s_pull[l]); }
        if (d->e_join[l]) cudaEventDestroy(d->e_join[l]);
    }
    if (d->e_fork) cudaEventDestroy(d->e_fork);
    for (int i = 0; i < kRing; i++)
        if (d->e_xready[i]) cudaEventDestroy(d->e_xready[i]);
    for (auto & b : d->blocks)
        if (b.owned && b.A) spmvb200_destroy(b.A);
    for (double * p : {d->X[0], d->X[1], d->Yh[0], d->Yh[1]})
        if (p) cudaFree(p);
    for (int i = 0; i < kRing; i++)
        for (cudaEvent_t ev : {d->e_exch[i], d->e_int[i], d->e_bnd[i], d->e_up[i], d->e_down[i]})
            if (ev) cudaEventDestroy(ev);
    if (d->t0) cudaEventDestroy(d->t0);
    if (d->t1) cudaEventDestroy(d->t1);
    for (cudaStream_t s : {d->s_int, d->s_bnd, d->s_comm, d->s_h2d, d->s_d2h})
        if (s) cudaStreamDestroy(s);
    delete d;
    return 0;
}

struct DistGuard {
    spmvb200_dist_t d;
    ~DistGuard() { if (d) dist_free(d); }
    spmvb200_dist_t release() { spmvb200_dist_t q = d; d = nullptr; return q; }
};

inline int slot(int64_t k) { return (int)(((k % kRing) + kRing) % kRing); }

// ---- peer-copy transport: host progress counters, handle exchange ------------------------------------------------
inline void publish(volatile int64_t * ctr, int64_t v)
{
    if (__atomic_load_n(ctr, __ATOMIC_RELAXED) < v) __atomic_store_n(ctr, v, __ATOMIC_RELEASE);
}

// Wait on the HOST until rank `who` has issued the record of the event we are about to wait on.
int spin_until(volatile int64_t * ctr, int64_t target, const char * what)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (int64_t it = 0; __atomic_load_n(ctr, __ATOMIC_ACQUIRE) < target; ++it) {
        if ((it & 1023) == 1023) {
            sched_yield();
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                return fail(SPMVB200_ERR_CUDA, std::string("peer-copy exchange: timed out waiting for a peer rank to issue ") + what);
        }
    }
    return 0;
}

struct PeerExport {
    cudaIpcMemHandle_t x[2];
    cudaIpcEventHandle_t xready[kRing], exch[kRing];
};

int setup_peer(spmvb200_dist_t d)
{
    NcclApi * nc = nccl_api();
    spmvb200_comm_t c = d->comm;
    const int P = d->P, rank = d->rank;
    // interprocess twins of the two event families the other ranks wait on
    for (int i = 0; i < kRing; i++) {
        if (d->e_exch[i]) cudaEventDestroy(d->e_exch[i]);
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_exch[i], cudaEventDisableTiming | cudaEventInterprocess));
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_xready[i], cudaEventDisableTiming | cudaEventInterprocess));
    }
    int lo_prio = 0, hi_prio = 0;
    SPMV_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    if (const char * env = getenv("SPMVB200_PULL_LANES")) d->pull_lanes = std::max(1, std::min(kPullStreams, atoi(env)));
    SPMV_CUDA(cudaEventCreateWithFlags(&d->e_fork, cudaEventDisableTiming));
    for (int l = 0; l < kPullStreams; l++) {
        SPMV_CUDA(cudaStreamCreateWithPriority(&d->s_pull[l], cudaStreamNonBlocking, hi_prio));
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_join[l], cudaEventDisableTiming));
    }
    PeerExport mine;
    memset(&mine, 0, sizeof mine);
    for (int b = 0; b < 2; b++) SPMV_CUDA(cudaIpcGetMemHandle(&mine.x[b], d->X[b]));
    for (int i = 0; i < kRing; i++) {
        SPMV_CUDA(cudaIpcGetEventHandle(&mine.xready[i], d->e_xready[i]));
        SPMV_CUDA(cudaIpcGetEventHandle(&mine.exch[i], d->e_exch[i]));
    }
    std::vector<PeerExport> all((size_t)P);
    char * dbuf = nullptr;
    SPMV_CUDA(cudaMalloc((void **)&dbuf, sizeof(PeerExport) * (size_t)P));
    SPMV_CUDA(cudaMemcpyAsync(dbuf + sizeof(PeerExport) * (size_t)rank, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = nc->AllGather(dbuf + sizeof(PeerExport) * (size_t)rank, dbuf, sizeof(PeerExport), ncclChar, c->nccl, c->stream);
    if (r != ncclSuccess) { cudaFree(dbuf); return nccl_fail(r, "ncclAllGather(peer handles)"); }
    SPMV_CUDA(cudaMemcpyAsync(all.data(), dbuf, sizeof(PeerExport) * (size_t)P, cudaMemcpyDeviceToHost, c->stream));
    SPMV_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(dbuf);
    for (int b = 0; b < 2; b++) d->peer_X[b].assign((size_t)P, nullptr);
    for (int i = 0; i < kRing; i++) { d->peer_xready[i].assign((size_t)P, nullptr); d->peer_exch[i].assign((size_t)P, nullptr); }
    // open only what this rank touches: the owners it pulls from (their buffers and readiness) and the ranks that pull
    // from it (their "pulls done" events)
    std::vector<char> pull_from((size_t)P, 0), pulled_by((size_t)P, 0);
    for (const Range & t : d->plan.recvs) pull_from[(size_t)t.peer] = 1;
    for (const Range & t : d->plan.sends) pulled_by[(size_t)t.peer] = 1;
    for (int q = 0; q < P; q++) {
        if (q == rank) {
            for (int b = 0; b < 2; b++) d->peer_X[b][(size_t)q] = d->X[b];
            continue;
        }
        if (pull_from[(size_t)q] || pulled_by[(size_t)q]) {  // (the push form writes into the buffers of the ranks that would pull)
            for (int b = 0; b < 2; b++)
                SPMV_CUDA(cudaIpcOpenMemHandle((void **)&d->peer_X[b][(size_t)q], all[(size_t)q].x[b], cudaIpcMemLazyEnablePeerAccess));
            for (int i = 0; i < kRing; i++) SPMV_CUDA(cudaIpcOpenEventHandle(&d->peer_xready[i][(size_t)q], all[(size_t)q].xready[i]));
        }
        if (pulled_by[(size_t)q])
            for (int i = 0; i < kRing; i++) SPMV_CUDA(cudaIpcOpenEventHandle(&d->peer_exch[i][(size_t)q], all[(size_t)q].exch[i]));
    }
    // the host progress counters: rank 0 creates the block, the others open it
    char name[64];
    snprintf(name, sizeof name, "/spmvb200_%016llx_%d", (unsigned long long)c->id_hash, c->executors);
    d->shm_bytes = sizeof(PeerCounters) * (size_t)P;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)d->shm_bytes) != 0) return fail(SPMVB200_ERR_IO, std::string("shm_open(") + name + ") failed");
    }
    SPMV_TRY(spmvb200_comm_barrier(c));  // the block exists (and is zero: ftruncate)
    if (rank != 0) {
        fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) return fail(SPMVB200_ERR_IO, std::string("shm_open(") + name + ") failed");
    }
    void * map = mmap(nullptr, d->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return fail(SPMVB200_ERR_IO, "mmap of the peer-copy progress block failed");
    d->shm = static_cast<PeerCounters *>(map);
    SPMV_TRY(spmvb200_comm_barrier(c));  // everybody has it mapped
    if (rank == 0) shm_unlink(name);
    d->peer = true;
    return 0;
}

// Fused compute + halo push (flag SPMVB200_DIST_PEER_PUSH, halo plan, peer transport): every send range of the plan is cut
// at the block boundaries; a block may push to at most two ranks and must run the sliced CSR kernel in store mode.  All
// ranks must agree (a receiver that expects a push must get one), hence the reduction at the end.
bool build_push_lists(spmvb200_dist_t d)
{
    bool ok = d->plan.mode == SPMVB200_EXCHANGE_HALO && !d->any_accumulate;
    d->push.assign(d->blocks.size(), {});
    for (size_t bi = 0; ok && bi < d->blocks.size(); bi++) {
        const DistBlock & b = d->blocks[bi];
        for (const Range & t : d->plan.sends) {
            const int64_t lo = std::max<int64_t>(t.lo - d->s, b.b), hi = std::min<int64_t>(t.hi - d->s, b.e);
            if (hi <= lo) continue;
            d->push[bi].push_back({t.peer, lo - b.b, hi - b.b});
        }
        if (d->push[bi].empty()) continue;
        if (d->push[bi].size() > 2 || !csr_uses_sliced_kernel(b.A) || b.A->opt_csr_threads != 0 || b.A->opt_csr_batch != 0) ok = false;
    }
    return ok;
}

int setup_push(spmvb200_dist_t d)
{
    double all_ok = (d->peer && build_push_lists(d)) ? 1.0 : 0.0;
    SPMV_TRY(spmvb200_comm_allreduce(d->comm, &all_ok, 2));
    d->push_active = all_ok > 0.5;
    if (!d->push_active) d->push.clear();
    return 0;
}

// ---- the exchange of step k (x_k lives in buffer `buf`), enqueued on d->s_comm ------------------------------------
// "The owner's slice of this x is complete" is e_int/e_bnd of step k-1 in the iteration (the kernels that produced
// it) and e_up of step k in run_host (the upload).  `lag`: how many steps ago the buffer was last in use (1 or 2).
enum class Ready { Iteration, Upload };

int wait_ready(cudaStream_t cs, spmvb200_dist_t o, Ready ready, int64_t k)
{
    if (ready == Ready::Iteration) {
        SPMV_CUDA(cudaStreamWaitEvent(cs, o->e_int[slot(k - 1)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(cs, o->e_bnd[slot(k - 1)], 0));
    } else {
        SPMV_CUDA(cudaStreamWaitEvent(cs, o->e_up[slot(k)], 0));
    }
    return 0;
}

int enqueue_exchange(spmvb200_dist_t d, int buf, int64_t k, Ready ready, int lag)
{
    cudaStream_t cs = d->s_comm;
    if (d->P == 1 || (d->plan.recvs.empty() && d->plan.sends.empty())) {
        SPMV_CUDA(cudaEventRecord(d->e_exch[slot(k)], cs));
        if (d->peer) publish(&d->shm[d->rank].exch_issued, k + 1);
        return 0;
    }
    // the remote ranges of this buffer were last read by this rank's remote blocks `lag` steps ago
    SPMV_CUDA(cudaStreamWaitEvent(cs, d->e_bnd[slot(k - lag)], 0));
    if (d->peer) {
        // One copy engine moves ~280 GB/s over NVLink; several pulls in flight on different streams use several engines.
        // The owners are visited starting behind this rank, so that at any time every owner serves different pullers.
        const size_t nr = d->plan.recvs.size();
        const int lanes = nr >= 3 ? d->pull_lanes : 1;
        size_t first = 0;
        while (first < nr && d->plan.recvs[first].peer < d->rank) first++;
        if (lanes > 1) {
            SPMV_CUDA(cudaEventRecord(d->e_fork, cs));  // carries the wait on e_bnd above to the pull streams
            for (int l = 0; l < lanes; l++) SPMV_CUDA(cudaStreamWaitEvent(d->s_pull[l], d->e_fork, 0));
        }
        for (size_t i = 0; i < nr; i++) {
            const Range & r = d->plan.recvs[(first + i) % nr];
            cudaStream_t ps = lanes > 1 ? d->s_pull[i % (size_t)lanes] : cs;
            // the owner has issued the record of "its slice of x_k is complete" (or set the slice synchronously)
            SPMV_TRY(spin_until(&d->shm[r.peer].xready_issued, k + 1, "the readiness of its slice of x"));
            SPMV_CUDA(cudaStreamWaitEvent(ps, d->peer_xready[slot(k)][(size_t)r.peer], 0));
            if (d->halo_by_push == k && ready == Ready::Iteration) continue;  // the owner's kernels stored it here already
            SPMV_CUDA(cudaMemcpyAsync(d->X[buf] + r.lo, d->peer_X[buf][(size_t)r.peer] + r.lo, sizeof(double) * (size_t)(r.hi - r.lo),
                                      cudaMemcpyDeviceToDevice, ps));
        }
        if (lanes > 1)
            for (int l = 0; l < lanes; l++) {
                SPMV_CUDA(cudaEventRecord(d->e_join[l], d->s_pull[l]));
                SPMV_CUDA(cudaStreamWaitEvent(cs, d->e_join[l], 0));
            }
    } else if (d->comm->local) {
        LocalGroup & g = *d->comm->group;
        for (const Range & r : d->plan.recvs) {
            spmvb200_dist_t o = nullptr;
            {
                std::lock_guard<std::mutex> lk(g.mu);
                o = g.dist[(size_t)r.peer];
            }
            if (!o) return fail(SPMVB200_ERR_INVALID, "in-process exchange: a peer rank has no executor");
            SPMV_TRY(wait_ready(cs, o, ready, k));
            if (ready == Ready::Iteration && o->push_active && o->halo_by_push == k) continue;  // the owner's kernels stored it here
            SPMV_CUDA(cudaMemcpyPeerAsync(d->X[buf] + r.lo, d->device, o->X[buf] + r.lo, o->device,
                                          sizeof(double) * (size_t)(r.hi - r.lo), cs));
        }
    } else {
        NcclApi * nc = nccl_api();
        SPMV_TRY(wait_ready(cs, d, ready, k));
        if (d->plan.mode == SPMVB200_EXCHANGE_ALLGATHER && d->equal_slices) {
            SPMV_NCCL(nc->AllGather(d->X[buf] + d->s, d->X[buf], (size_t)d->rows, ncclDouble, d->comm->nccl, cs));  // in place
        } else if (d->plan.mode == SPMVB200_EXCHANGE_ALLGATHER) {
            SPMV_NCCL(nc->GroupStart());
            for (int q = 0; q < d->P; q++) {
                const int64_t a = d->starts[(size_t)q], b = d->starts[(size_t)q + 1];
                if (b > a) SPMV_NCCL(nc->Broadcast(d->X[buf] + a, d->X[buf] + a, (size_t)(b - a), ncclDouble, q, d->comm->nccl, cs));
            }
            SPMV_NCCL(nc->GroupEnd());
        } else {
            SPMV_NCCL(nc->GroupStart());
            for (const Range & r : d->plan.sends)
                SPMV_NCCL(nc->Send(d->X[buf] + r.lo, (size_t)(r.hi - r.lo), ncclDouble, r.peer, d->comm->nccl, cs));
            for (const Range & r : d->plan.recvs)
                SPMV_NCCL(nc->Recv(d->X[buf] + r.lo, (size_t)(r.hi - r.lo), ncclDouble, r.peer, d->comm->nccl, cs));
            SPMV_NCCL(nc->GroupEnd());
        }
    }
    SPMV_CUDA(cudaEventRecord(d->e_exch[slot(k)], cs));
    if (d->peer) publish(&d->shm[d->rank].exch_issued, k + 1);
    return 0;
}

// Before a step overwrites the rank's slice of a buffer: whoever was still reading that slice must be done.
// The previous occupant of the buffer written at step k is x_(k-1); its readers outside this rank's own compute
// streams are the exchange of step k-1: this rank's sends (NCCL) or the peers' pulls (in-process).
int wait_slice_readers(spmvb200_dist_t d, cudaStream_t s, int64_t k_prev_exchange)
{
    if (d->P == 1) return 0;
    const int sl = slot(k_prev_exchange);
    if (d->peer) {
        if (k_prev_exchange < 0) return 0;
        int last = -1;
        for (const Range & r : d->plan.sends) {  // the ranks that pull from this one
            if (r.peer == last) continue;
            last = r.peer;
            SPMV_TRY(spin_until(&d->shm[r.peer].exch_issued, k_prev_exchange + 1, "its pulls of the previous step"));
            SPMV_CUDA(cudaStreamWaitEvent(s, d->peer_exch[sl][(size_t)r.peer], 0));
        }
        return 0;
    }
    if (!d->comm->local) {
        SPMV_CUDA(cudaStreamWaitEvent(s, d->e_exch[sl], 0));
        return 0;
    }
    LocalGroup & g = *d->comm->group;
    int last = -1;
    for (const Range & r : d->plan.sends) {
        if (r.peer == last) continue;
        last = r.peer;
        spmvb200_dist_t p = nullptr;
        {
            std::lock_guard<std::mutex> lk(g.mu);
            p = g.dist[(size_t)r.peer];
        }
        if (p) SPMV_CUDA(cudaStreamWaitEvent(s, p->e_exch[sl], 0));
    }
    return 0;
}

// push_buf >= 0: iteration step k whose results also go to the neighbours' buffer `push_buf` (their X[(k+1) % 2]).
int launch_blocks(spmvb200_dist_t d, bool remote, const double * x, double * y_base, double alpha, int push_buf = -1,
                  int64_t k = 0)
{
    for (size_t bi = 0; bi < d->blocks.size(); bi++) {
        DistBlock & b = d->blocks[bi];
        if (b.remote != remote) continue;
        SPMV_TRY(spmvb200_bind_x(b.A, (void *)x));
        SPMV_TRY(spmvb200_bind_y(b.A, (void *)(y_base + b.b)));
        SPMV_TRY(spmvb200_set_alpha(b.A, alpha));
        for (int t = 0; t < 2; t++) { b.A->push_y[t] = nullptr; b.A->push_lo[t] = b.A->push_hi[t] = 0; }
        if (push_buf >= 0 && d->push_active) {
            int t = 0;
            for (const auto & pr : d->push[bi]) {
                // the target's buffer still holds x_(k-1), which its boundary kernels of step k-1 read: they are behind
                // its "x_k is ready" event
                cudaStream_t bs = remote ? d->s_bnd : d->s_int;
                double * target = nullptr;
                if (d->comm->local) {
                    spmvb200_dist_t q = nullptr;
                    {
                        std::lock_guard<std::mutex> lk(d->comm->group->mu);
                        q = d->comm->group->dist[(size_t)pr.peer];
                    }
                    if (!q) return fail(SPMVB200_ERR_INVALID, "in-process halo push: a peer rank has no executor");
                    SPMV_TRY(wait_ready(bs, q, Ready::Iteration, k));
                    target = q->X[push_buf];
                } else {
                    SPMV_TRY(spin_until(&d->shm[pr.peer].xready_issued, k + 1, "its previous step (halo push)"));
                    SPMV_CUDA(cudaStreamWaitEvent(bs, d->peer_xready[slot(k)][(size_t)pr.peer], 0));
                    target = d->peer_X[push_buf][(size_t)pr.peer];
                }
                b.A->push_y[t] = target + d->s + b.b;  // same global index in the peer's buffer
                b.A->push_lo[t] = pr.lo;
                b.A->push_hi[t] = pr.hi;
                t++;
            }
        }
        const int rc = spmvb200_spmv(b.A);
        for (int t = 0; t < 2; t++) { b.A->push_y[t] = nullptr; b.A->push_lo[t] = b.A->push_hi[t] = 0; }
        SPMV_TRY(rc);
    }
    return 0;
}

int ensure_plan(spmvb200_dist_t d)
{
    if (d->plan_ready) return 0;
    if (!d->comm->local) return fail(SPMVB200_ERR_INVALID, "executor has no exchange plan");
    LocalGroup & g = *d->comm->group;
    std::lock_guard<std::mutex> lk(g.mu);
    for (int q = 0; q < g.nranks; q++)
        if (!g.need_set[(size_t)q]) return fail(SPMVB200_ERR_INVALID, "in-process communicator: create the executor of every rank before the first step");
    if (g.wanted_mode == SPMVB200_EXCHANGE_ALLGATHER) d->wanted_mode = SPMVB200_EXCHANGE_ALLGATHER;  // one plan for all ranks
    finalize_plan(d, g.need_lo.data(), g.need_hi.data());
    if (d->want_push) {
        d->push_active = build_push_lists(d);
        if (!d->push_active) d->push.clear();
    }
    return 0;
}

// One iteration step.
int step(spmvb200_dist_t d, double alpha)
{
    SPMV_TRY(ensure_plan(d));
    const int64_t k = d->k;
    const int cur = (int)(k & 1), nxt = cur ^ 1;
    SPMV_TRY(enqueue_exchange(d, cur, k, Ready::Iteration, 1));
    // blocks that need no remote x
    SPMV_CUDA(cudaStreamWaitEvent(d->s_int, d->e_bnd[slot(k - 1)], 0));
    SPMV_TRY(wait_slice_readers(d, d->s_int, k - 1));
    const int push_buf = d->push_active ? nxt : -1;
    SPMV_TRY(launch_blocks(d, false, d->X[cur], d->X[nxt] + d->s, alpha, push_buf, k));
    SPMV_CUDA(cudaEventRecord(d->e_int[slot(k)], d->s_int));
    // blocks that do
    SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_exch[slot(k)], 0));
    SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_int[slot(k - 1)], 0));
    SPMV_TRY(wait_slice_readers(d, d->s_bnd, k - 1));
    if (d->any_accumulate) SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_int[slot(k)], 0));
    SPMV_TRY(launch_blocks(d, true, d->X[cur], d->X[nxt] + d->s, alpha, push_buf, k));
    SPMV_CUDA(cudaEventRecord(d->e_bnd[slot(k)], d->s_bnd));
    if (d->push_active) d->halo_by_push = k + 1;  // the neighbours hold their halo of x_(k+1) once their wait on e_xready passes
    if (d->peer) {  // "this rank's slice of x_(k+1) is complete", for the ranks that will pull from it
        SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_int[slot(k)], 0));
        SPMV_CUDA(cudaEventRecord(d->e_xready[slot(k + 1)], d->s_bnd));
        publish(&d->shm[d->rank].xready_issued, k + 2);
    }
    d->k = k + 1;
    return 0;
}

int sync_all(spmvb200_dist_t d)
{
    for (cudaStream_t s : {d->s_comm, d->s_int, d->s_bnd, d->s_h2d, d->s_d2h})
        if (s) SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace

extern "C" {

int spmvb200_exchange_plan(int32_t parts, const int64_t * starts, const int64_t * need_lo, const int64_t * need_hi,
                           int32_t rank, int32_t mode, int32_t cap, int32_t * chosen_mode, int32_t * n_sends,
                           int64_t * sends, int32_t * n_recvs, int64_t * recvs, int64_t * recv_bytes)
try {
    if (parts < 1 || !starts || !need_lo || !need_hi || rank < 0 || rank >= parts || mode < 0 || mode > 2 || cap < 0)
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    const Plan p = make_plan(parts, starts, need_lo, need_hi, rank, mode);
    if (chosen_mode) *chosen_mode = p.mode;
    if (n_sends) *n_sends = (int32_t)p.sends.size();
    if (n_recvs) *n_recvs = (int32_t)p.recvs.size();
    if (recv_bytes) *recv_bytes = p.recv_bytes;
    for (size_t i = 0; sends && i < p.sends.size() && (int32_t)i < cap; i++) {
        sends[3 * i] = p.sends[i].peer; sends[3 * i + 1] = p.sends[i].lo; sends[3 * i + 2] = p.sends[i].hi;
    }
    for (size_t i = 0; recvs && i < p.recvs.size() && (int32_t)i < cap; i++) {
        recvs[3 * i] = p.recvs[i].peer; recvs[3 * i + 1] = p.recvs[i].lo; recvs[3 * i + 2] = p.recvs[i].hi;
    }
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_comm_create_local(int nranks, const int * devices, spmvb200_comm_t * comms)
try {
    if (nranks < 1 || !comms) return fail(SPMVB200_ERR_INVALID, "bad argument");
    int ndev = 0;
    SPMV_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(SPMVB200_ERR_CUDA, "no CUDA device");
    auto group = std::make_shared<LocalGroup>();
    group->nranks = nranks;
    group->dist.assign((size_t)nranks, nullptr);
    group->need_lo.assign((size_t)nranks, 0);
    group->need_hi.assign((size_t)nranks, 0);
    group->need_set.assign((size_t)nranks, 0);
    group->vals.assign((size_t)nranks, 0.0);
    int keep = 0;
    cudaGetDevice(&keep);
    for (int r = 0; r < nranks; r++) comms[r] = nullptr;
    for (int r = 0; r < nranks; r++) {
        spmvb200_comm_t c = new spmvb200_comm_s();
        c->rank = r; c->nranks = nranks; c->local = true; c->group = group;
        c->device = devices ? devices[r] : r % ndev;
        comms[r] = c;
        if (c->device < 0 || c->device >= ndev) {
            for (int q = 0; q <= r; q++) { delete comms[q]; comms[q] = nullptr; }
            return fail(SPMVB200_ERR_INVALID, "bad device index");
        }
    }
    // direct peer access between the ranks' devices (NVLink); without it the peer copies are staged by the driver
    for (int r = 0; r < nranks; r++) {
        cudaSetDevice(comms[r]->device);
        for (int q = 0; q < nranks; q++) {
            if (comms[q]->device == comms[r]->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, comms[r]->device, comms[q]->device) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(comms[q]->device, 0);
                if (e != cudaSuccess) cudaGetLastError();  // already enabled
            }
        }
    }
    cudaSetDevice(keep);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_comm_unique_id(void * id)
try {
    if (!id) return fail(SPMVB200_ERR_INVALID, "null argument");
    NcclApi * nc = nccl_api();
    if (!nc->handle) return fail(SPMVB200_ERR_UNSUPPORTED, nc->error);
    static_assert(sizeof(ncclUniqueId) == SPMVB200_COMM_ID_BYTES, "id size");
    ncclUniqueId u;
    SPMV_NCCL(nc->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_comm_create_nccl(const void * id, int rank, int nranks, spmvb200_comm_t * out)
try {
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(SPMVB200_ERR_INVALID, "bad argument");
    NcclApi * nc = nccl_api();
    if (!nc->handle) return fail(SPMVB200_ERR_UNSUPPORTED, nc->error);
    std::unique_ptr<spmvb200_comm_s> c(new spmvb200_comm_s());
    c->rank = rank; c->nranks = nranks; c->local = false;
    SPMV_CUDA(cudaGetDevice(&c->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    for (size_t i = 0; i < sizeof u; i++) c->id_hash = c->id_hash * 1099511628211ull + (unsigned char)u.internal[i] + 1;
    SPMV_NCCL(nc->CommInitRank(&c->nccl, nranks, u, rank));
    SPMV_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SPMV_CUDA(cudaMalloc((void **)&c->scratch, 64 * sizeof(double)));
    *out = c.release();
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_comm_rank(spmvb200_comm_t c, int * rank, int * nranks, int * device)
try {
    if (!c) return fail(SPMVB200_ERR_INVALID, "null communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (device) *device = c->device;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_comm_allreduce(spmvb200_comm_t c, double * value, int op)
try {
    SPMV_TRY(comm_check(c));
    if (!value || op < 0 || op > 2) return fail(SPMVB200_ERR_INVALID, "bad argument");
    if (c->nranks == 1) return 0;
    if (c->local) {
        *value = group_reduce(*c->group, c->rank, *value, op);
        return 0;
    }
    NcclApi * nc = nccl_api();
    SPMV_CUDA(cudaMemcpyAsync(c->scratch, value, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SPMV_NCCL(nc->AllReduce(c->scratch, c->scratch, 1, ncclDouble, op == 0 ? ncclMax : op == 1 ? ncclSum : ncclMin, c->nccl, c->stream));
    SPMV_CUDA(cudaMemcpyAsync(value, c->scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SPMV_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_comm_barrier(spmvb200_comm_t c)
try {
    double v = 0.0;
    return spmvb200_comm_allreduce(c, &v, 1);
}
SPMV_ABI_CATCH

int spmvb200_comm_destroy(spmvb200_comm_t c)
try {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->nccl) nccl_api()->CommDestroy(c->nccl);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->scratch) cudaFree(c->scratch);
    delete c;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_create(spmvb200_comm_t comm, spmvb200_matrix_t local, const int64_t * starts, int32_t exchange,
                         int32_t format, int32_t flags, spmvb200_dist_t * out)
try {
    SPMV_TRY(comm_check(comm));
    if (!local || !starts || !out || exchange < 0 || exchange > 2 || format < 0 || format > 3)
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    *out = nullptr;
    const int P = comm->nranks, rank = comm->rank;
    if (starts[0] != 0) return fail(SPMVB200_ERR_INVALID, "starts[0] must be 0");
    for (int q = 0; q < P; q++)
        if (starts[q] > starts[q + 1]) return fail(SPMVB200_ERR_INVALID, "row starts must be non-decreasing");
    if (local->device != comm->device) return fail(SPMVB200_ERR_INVALID, "the row block lives on another device than the communicator's rank");
    spmvb200_info inf;
    SPMV_TRY(spmvb200_matrix_info(local, &inf));
    DistGuard guard{new spmvb200_dist_s()};
    spmvb200_dist_t d = guard.d;
    d->comm = comm; d->rank = rank; d->P = P; d->device = comm->device;
    d->starts.assign(starts, starts + P + 1);
    d->n = starts[P]; d->s = starts[rank]; d->e = starts[rank + 1]; d->rows = d->e - d->s;
    d->nnz = inf.num_entries;
    if (inf.rows != d->rows) return fail(SPMVB200_ERR_INVALID, "the row block does not have starts[rank+1]-starts[rank] rows");
    if (inf.columns != d->n) return fail(SPMVB200_ERR_INVALID, "the row block must keep global column indices (columns == starts[nranks])");
    d->equal_slices = true;
    for (int q = 0; q < P; q++) d->equal_slices = d->equal_slices && (starts[q + 1] - starts[q] == d->rows);

    int lo_prio = 0, hi_prio = 0;
    SPMV_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    SPMV_CUDA(cudaStreamCreateWithPriority(&d->s_int, cudaStreamNonBlocking, lo_prio));
    SPMV_CUDA(cudaStreamCreateWithPriority(&d->s_bnd, cudaStreamNonBlocking, hi_prio));
    SPMV_CUDA(cudaStreamCreateWithPriority(&d->s_comm, cudaStreamNonBlocking, hi_prio));
    for (int i = 0; i < kRing; i++) {
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_exch[i], cudaEventDisableTiming));
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_int[i], cudaEventDisableTiming));
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_bnd[i], cudaEventDisableTiming));
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_up[i], cudaEventDisableTiming));
        SPMV_CUDA(cudaEventCreateWithFlags(&d->e_down[i], cudaEventDisableTiming));
    }
    SPMV_CUDA(cudaEventCreate(&d->t0));
    SPMV_CUDA(cudaEventCreate(&d->t1));
    for (int b = 0; b < 2; b++) {
        SPMV_CUDA(cudaMalloc((void **)&d->X[b], sizeof(double) * (size_t)(d->n + 16)));
        SPMV_CUDA(cudaMemset(d->X[b], 0, sizeof(double) * (size_t)(d->n + 16)));
    }
    d->x_bytes = 2 * 8 * (d->n + 16);

    // ---- what this rank's rows reference, and how they are cut into blocks ------------------------------------------
    const bool csr = inf.format == SPMVB200_CSR;
    const bool column_split = (flags & SPMVB200_DIST_COLUMN_SPLIT) && csr && P > 1;
    const bool overlap = !(flags & SPMVB200_DIST_NO_OVERLAP) && P > 1;
    int64_t col_min = 0, col_max = d->n - 1, lo_end = d->rows, hi_begin = 0;
    if (csr && !column_split) SPMV_TRY(spmvb200_csr_column_span(local, d->s, d->e, &col_min, &col_max, &lo_end, &hi_begin));
    else if (P > 1) exchange = SPMVB200_EXCHANGE_ALLGATHER;  // no column analysis: the rank is taken to need all of x
    d->need_lo = col_max < 0 ? 0 : col_min;
    d->need_hi = col_max < 0 ? 0 : col_max + 1;
    d->wanted_mode = exchange;
    d->want_push = (flags & SPMVB200_DIST_PEER_PUSH) != 0 && P > 1;
    auto add = [&](spmvb200_matrix_t A, int64_t b, int64_t e, bool remote, bool accumulate, bool owned) -> int {
        if (format != SPMVB200_CSR) {
            spmvb200_info bi;
            SPMV_TRY(spmvb200_matrix_info(A, &bi));
            if (bi.format == SPMVB200_CSR) {
                spmvb200_matrix_t conv = nullptr;
                SPMV_TRY(spmvb200_convert(A, format, 0, &conv));
                if (owned) spmvb200_destroy(A);
                A = conv; owned = true;
            }
        }
        DistBlock blk;
        blk.A = A; blk.b = b; blk.e = e; blk.remote = remote; blk.accumulate = accumulate; blk.owned = owned;
        d->blocks.push_back(blk);
        return 0;
    };
    const bool consume = (flags & SPMVB200_DIST_CONSUME_LOCAL) != 0;
    if (column_split) {
        spmvb200_matrix_t inside = nullptr, outside = nullptr;
        SPMV_TRY(spmvb200_csr_column_split(local, d->s, d->e, &inside, &outside));
        int rc = add(inside, 0, d->rows, false, false, true);
        if (rc) { spmvb200_destroy(outside); return rc; }
        SPMV_TRY(add(outside, 0, d->rows, true, true, true));
        d->any_accumulate = true;
    } else if (!overlap || !csr || lo_end >= hi_begin) {
        SPMV_TRY(add(local, 0, d->rows, P > 1, false, false));
    } else {
        struct Cut { int64_t b, e; bool remote; };
        std::vector<Cut> cuts;
        if (lo_end > 0) cuts.push_back({0, lo_end, true});
        cuts.push_back({lo_end, hi_begin, false});
        if (hi_begin < d->rows) cuts.push_back({hi_begin, d->rows, true});
        for (auto & c : cuts) {
            if (c.b == 0 && c.e == d->rows) {
                SPMV_TRY(add(local, 0, d->rows, c.remote, false, false));
            } else {
                spmvb200_matrix_t blk = nullptr;
                SPMV_TRY(spmvb200_csr_row_block(local, c.b, c.e, &blk));
                SPMV_TRY(add(blk, c.b, c.e, c.remote, false, true));
            }
        }
    }
    for (auto & b : d->blocks) {
        SPMV_TRY(spmvb200_set_stream(b.A, b.remote ? d->s_bnd : d->s_int));
        SPMV_TRY(spmvb200_set_option(b.A, "beta0", b.accumulate ? 0 : 1));
        SPMV_TRY(spmvb200_set_option(b.A, "csr.drop_row_major", 1));  // a rank keeps one copy of its entries
        SPMV_TRY(spmvb200_prepare(b.A));
    }

    // ---- the exchange plan needs every rank's column range ---------------------------------------------------------
    if (P == 1) {
        finalize_plan(d, &d->need_lo, &d->need_hi);
    } else if (comm->local) {
        LocalGroup & g = *comm->group;
        std::lock_guard<std::mutex> lk(g.mu);
        if (g.dist[(size_t)rank]) return fail(SPMVB200_ERR_INVALID, "this rank of the communicator already has an executor");
        g.dist[(size_t)rank] = d;
        g.need_lo[(size_t)rank] = d->need_lo;
        g.need_hi[(size_t)rank] = d->need_hi;
        g.need_set[(size_t)rank] = 1;
        if (d->wanted_mode == SPMVB200_EXCHANGE_ALLGATHER || g.wanted_mode < 0) g.wanted_mode = d->wanted_mode;
    } else {
        NcclApi * nc = nccl_api();
        std::vector<int64_t> all((size_t)3 * P);
        int64_t * dbuf = reinterpret_cast<int64_t *>(comm->scratch);  // 64 doubles: up to 21 ranks
        if (3 * P > 64) return fail(SPMVB200_ERR_UNSUPPORTED, "more than 21 ranks");
        const int64_t mine[3] = {d->need_lo, d->need_hi, d->wanted_mode};
        SPMV_CUDA(cudaMemcpyAsync(dbuf + 3 * rank, mine, sizeof mine, cudaMemcpyHostToDevice, comm->stream));
        SPMV_NCCL(nc->AllGather(dbuf + 3 * rank, dbuf, 3, ncclInt64, comm->nccl, comm->stream));
        SPMV_CUDA(cudaMemcpyAsync(all.data(), dbuf, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, comm->stream));
        SPMV_CUDA(cudaStreamSynchronize(comm->stream));
        std::vector<int64_t> lo((size_t)P), hi((size_t)P);
        for (int q = 0; q < P; q++) {
            lo[(size_t)q] = all[(size_t)3 * q]; hi[(size_t)q] = all[(size_t)3 * q + 1];
            // every rank must run the same plan: a rank whose block allows no column analysis (all-gather) decides for all
            if (all[(size_t)3 * q + 2] == SPMVB200_EXCHANGE_ALLGATHER) d->wanted_mode = SPMVB200_EXCHANGE_ALLGATHER;
        }
        for (int q = 0; q < P; q++)
            if (d->wanted_mode != SPMVB200_EXCHANGE_ALLGATHER && all[(size_t)3 * q + 2] != d->wanted_mode)
                return fail(SPMVB200_ERR_INVALID, "the ranks ask for different exchange modes");
        finalize_plan(d, lo.data(), hi.data());
        if (flags & (SPMVB200_DIST_PEER_COPY | SPMVB200_DIST_PEER_PUSH)) SPMV_TRY(setup_peer(d));
        if (flags & SPMVB200_DIST_PEER_PUSH) SPMV_TRY(setup_push(d));
        comm->executors++;
    }
    // `local` changes hands only now that nothing can fail any more: until here a failure leaves it with the caller
    if (consume) {
        bool used = false;
        for (auto & b : d->blocks)
            if (b.A == local) { b.owned = true; used = true; }
        if (!used) spmvb200_destroy(local);
    }
    *out = guard.release();
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_set_x(spmvb200_dist_t d, const double * x)
try {
    SPMV_TRY(dist_check(d));
    if (!x && d->rows > 0) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_TRY(sync_all(d));
    if (d->rows > 0)
        SPMV_CUDA(cudaMemcpy(d->X[d->k & 1] + d->s, x, sizeof(double) * (size_t)d->rows, cudaMemcpyHostToDevice));
    d->halo_by_push = -1;  // this x did not come out of the kernels: the next exchange copies
    if (d->peer) {
        // the other ranks pull this slice without asking: nobody may proceed before every slice is in place
        // (spmvb200_dist_set_x is collective with the peer-copy transport)
        publish(&d->shm[d->rank].xready_issued, d->k + 1);
        SPMV_TRY(spmvb200_comm_barrier(d->comm));
    }
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_get_x(spmvb200_dist_t d, double * x)
try {
    SPMV_TRY(dist_check(d));
    if (!x && d->rows > 0) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_TRY(sync_all(d));
    if (d->rows > 0)
        SPMV_CUDA(cudaMemcpy(x, d->X[d->k & 1] + d->s, sizeof(double) * (size_t)d->rows, cudaMemcpyDeviceToHost));
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_x_device(spmvb200_dist_t d, void ** ptr)
try {
    if (!d || !ptr) return fail(SPMVB200_ERR_INVALID, "null argument");
    *ptr = d->X[d->k & 1] + d->s;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_step(spmvb200_dist_t d, double alpha)
try {
    SPMV_TRY(dist_check(d));
    return step(d, alpha);
}
SPMV_ABI_CATCH

int spmvb200_dist_sync(spmvb200_dist_t d)
try {
    SPMV_TRY(dist_check(d));
    return sync_all(d);
}
SPMV_ABI_CATCH

int spmvb200_dist_time(const spmvb200_dist_t * ds, int n, int warmup, int steps, double alpha, float * ms)
try {
    if (!ds || n < 1 || warmup < 0 || steps < 1 || !ms) return fail(SPMVB200_ERR_INVALID, "bad argument");
    for (int r = 0; r < n; r++) {
        SPMV_TRY(dist_check(ds[r]));
        if (!ds[r]->comm->local && n != 1) return fail(SPMVB200_ERR_INVALID, "an NCCL rank is driven by its own process: n must be 1");
        if (ds[r]->comm->local && n != ds[r]->P) return fail(SPMVB200_ERR_INVALID, "pass the executors of ALL ranks of the in-process communicator");
    }
    for (int w = 0; w < warmup; w++)
        for (int r = 0; r < n; r++) { SPMV_TRY(dist_check(ds[r])); SPMV_TRY(step(ds[r], alpha)); }
    for (int r = 0; r < n; r++) { SPMV_TRY(dist_check(ds[r])); SPMV_TRY(sync_all(ds[r])); }
    if (!ds[0]->comm->local && ds[0]->P > 1) SPMV_TRY(spmvb200_comm_barrier(ds[0]->comm));
    for (int r = 0; r < n; r++) {
        spmvb200_dist_t d = ds[r];
        SPMV_TRY(dist_check(d));
        SPMV_CUDA(cudaEventRecord(d->t0, d->s_int));
        // the other streams start their part of the first timed step after t0
        SPMV_CUDA(cudaStreamWaitEvent(d->s_comm, d->t0, 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->t0, 0));
    }
    for (int k = 0; k < steps; k++)
        for (int r = 0; r < n; r++) { SPMV_TRY(dist_check(ds[r])); SPMV_TRY(step(ds[r], alpha)); }
    for (int r = 0; r < n; r++) {
        spmvb200_dist_t d = ds[r];
        SPMV_TRY(dist_check(d));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_int, d->e_bnd[slot(d->k - 1)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_int, d->e_exch[slot(d->k - 1)], 0));
        SPMV_CUDA(cudaEventRecord(d->t1, d->s_int));
    }
    for (int r = 0; r < n; r++) {
        spmvb200_dist_t d = ds[r];
        SPMV_TRY(dist_check(d));
        SPMV_CUDA(cudaEventSynchronize(d->t1));
        SPMV_CUDA(cudaEventElapsedTime(&ms[r], d->t0, d->t1));
        SPMV_TRY(sync_all(d));
    }
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_run_host(spmvb200_dist_t d, int steps, const double * const * xs, double * const * ys, double alpha,
                           float * ms)
try {
    SPMV_TRY(dist_check(d));
    if (steps < 1 || !xs || !ys) return fail(SPMVB200_ERR_INVALID, "bad argument");
    SPMV_TRY(ensure_plan(d));
    if (!d->s_h2d) {
        SPMV_CUDA(cudaStreamCreateWithFlags(&d->s_h2d, cudaStreamNonBlocking));
        SPMV_CUDA(cudaStreamCreateWithFlags(&d->s_d2h, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            SPMV_CUDA(cudaMalloc((void **)&d->Yh[b], sizeof(double) * (size_t)(d->rows + 16)));
            SPMV_CUDA(cudaMemset(d->Yh[b], 0, sizeof(double) * (size_t)(d->rows + 16)));
        }
    }
    SPMV_TRY(sync_all(d));
    d->halo_by_push = -1;
    const bool threads_meet = d->comm->local && d->P > 1;  // one host thread per rank: they meet once per step
    if (d->P > 1) SPMV_TRY(spmvb200_comm_barrier(d->comm));
    SPMV_CUDA(cudaEventRecord(d->t0, d->s_h2d));
    const size_t bytes = sizeof(double) * (size_t)d->rows;
    for (int i = 0; i < steps; i++) {
        const int64_t k = d->k;  // the event slots keep counting across iteration steps and host steps
        const int buf = (int)(k & 1);
        // upload of x_i's slice: the buffer's previous occupant (two steps ago) must have been consumed by this rank's
        // kernels and by the exchange that read it
        SPMV_CUDA(cudaStreamWaitEvent(d->s_h2d, d->e_int[slot(k - 2)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_h2d, d->e_bnd[slot(k - 2)], 0));
        SPMV_TRY(wait_slice_readers(d, d->s_h2d, k - 2));
        if (d->rows > 0) SPMV_CUDA(cudaMemcpyAsync(d->X[buf] + d->s, xs[i], bytes, cudaMemcpyHostToDevice, d->s_h2d));
        SPMV_CUDA(cudaEventRecord(d->e_up[slot(k)], d->s_h2d));
        if (d->peer) {
            SPMV_CUDA(cudaEventRecord(d->e_xready[slot(k)], d->s_h2d));
            publish(&d->shm[d->rank].xready_issued, k + 1);
        }
        if (threads_meet) group_reduce(*d->comm->group, d->rank, 0.0, 1);  // every rank's e_up of this step is recorded
        SPMV_TRY(enqueue_exchange(d, buf, k, Ready::Upload, 2));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_int, d->e_up[slot(k)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_int, d->e_down[slot(k - 2)], 0));
        SPMV_TRY(launch_blocks(d, false, d->X[buf], d->Yh[buf], alpha));
        SPMV_CUDA(cudaEventRecord(d->e_int[slot(k)], d->s_int));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_exch[slot(k)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_up[slot(k)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_down[slot(k - 2)], 0));
        if (d->any_accumulate) SPMV_CUDA(cudaStreamWaitEvent(d->s_bnd, d->e_int[slot(k)], 0));
        SPMV_TRY(launch_blocks(d, true, d->X[buf], d->Yh[buf], alpha));
        SPMV_CUDA(cudaEventRecord(d->e_bnd[slot(k)], d->s_bnd));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_d2h, d->e_int[slot(k)], 0));
        SPMV_CUDA(cudaStreamWaitEvent(d->s_d2h, d->e_bnd[slot(k)], 0));
        if (d->rows > 0) SPMV_CUDA(cudaMemcpyAsync(ys[i], d->Yh[buf], bytes, cudaMemcpyDeviceToHost, d->s_d2h));
        SPMV_CUDA(cudaEventRecord(d->e_down[slot(k)], d->s_d2h));
        d->k = k + 1;
    }
    SPMV_CUDA(cudaStreamWaitEvent(d->s_d2h, d->e_exch[slot(d->k - 1)], 0));
    SPMV_CUDA(cudaEventRecord(d->t1, d->s_d2h));
    SPMV_CUDA(cudaEventSynchronize(d->t1));
    if (ms) SPMV_CUDA(cudaEventElapsedTime(ms, d->t0, d->t1));
    return sync_all(d);
}
SPMV_ABI_CATCH

int spmvb200_dist_info(spmvb200_dist_t d, spmvb200_dist_info_t * info)
try {
    if (!d || !info) return fail(SPMVB200_ERR_INVALID, "null argument");
    memset(info, 0, sizeof *info);
    if (!d->plan_ready && d->comm->local) ensure_plan(d);  // may still be incomplete: the counts then read 0
    info->rank = d->rank; info->nranks = d->P;
    info->exchange = d->plan_ready ? d->plan.mode : d->wanted_mode;
    info->n_blocks = (int32_t)d->blocks.size();
    info->n_sends = (int32_t)d->plan.sends.size(); info->n_recvs = (int32_t)d->plan.recvs.size();
    info->recv_bytes_per_step = d->plan.recv_bytes; info->send_bytes_per_step = d->plan.send_bytes;
    info->rows = d->rows; info->row_begin = d->s; info->num_entries = d->nnz;
    int64_t dev = d->x_bytes, launches = 0;
    for (auto & b : d->blocks) {
        spmvb200_info bi;
        if (spmvb200_matrix_info(b.A, &bi) == 0) {
            dev += bi.device_bytes;
            launches += bi.format == SPMVB200_HYB ? 2 : 1;
        }
        if (!b.remote) info->interior_rows += b.e - b.b;
    }
    info->device_bytes = dev;
    info->launches_per_step = launches;
    info->steps_done = d->k;
    info->halo_push = d->push_active ? 1 : 0;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_block(spmvb200_dist_t d, int32_t b, int64_t * row_begin, int64_t * row_end, int32_t * needs_remote_x,
                        spmvb200_matrix_t * matrix)
try {
    if (!d || b < 0 || (size_t)b >= d->blocks.size()) return fail(SPMVB200_ERR_INVALID, "bad argument");
    const DistBlock & k = d->blocks[(size_t)b];
    if (row_begin) *row_begin = k.b;
    if (row_end) *row_end = k.e;
    if (needs_remote_x) *needs_remote_x = k.remote ? 1 : 0;
    if (matrix) *matrix = k.A;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_dist_destroy(spmvb200_dist_t d)
try {
    return dist_free(d);
}
SPMV_ABI_CATCH

}  // extern "C"
