// comm.cu -- NCCL binding (run-time dlopen) for the row-sharded multi-GPU paths.
#include "../../include/b200admm.h"
#include "comm.h"
#include <dlfcn.h>
#include <cstring>
#include <vector>

namespace b200 {

namespace {

// the handful of NCCL entry points we use, with the ABI of nccl.h 2.27 / 2.28
typedef struct { char internal[128]; } NcclId;
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int /*dtype*/, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
constexpr int NCCL_F32 = 7, NCCL_F64 = 8, NCCL_SUM = 0, NCCL_U8 = 1;

struct NcclApi {
    void* lib = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_allgather allgather = nullptr;
    fn_errstr errstr = nullptr;
};

NcclApi& api()
{
    static NcclApi a;
    if (!a.lib) {
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) throw CodeError(B200ADMM_ENCCL, std::string("cannot load NCCL: ") + dlerror());
        a.get_id = (fn_get_id)dlsym(a.lib, "ncclGetUniqueId");
        a.init_rank = (fn_init_rank)dlsym(a.lib, "ncclCommInitRank");
        a.destroy = (fn_destroy)dlsym(a.lib, "ncclCommDestroy");
        a.allreduce = (fn_allreduce)dlsym(a.lib, "ncclAllReduce");
        a.allgather = (fn_allgather)dlsym(a.lib, "ncclAllGather");
        a.errstr = (fn_errstr)dlsym(a.lib, "ncclGetErrorString");
        if (!a.get_id || !a.init_rank || !a.destroy || !a.allreduce)
            throw CodeError(B200ADMM_ENCCL, "NCCL library lacks a required symbol");
    }
    return a;
}

void nccl_check(int rc, const char* what)
{
    if (rc != 0) {
        NcclApi& a = api();
        throw CodeError(B200ADMM_ENCCL, std::string(what) + " failed: " + (a.errstr ? a.errstr(rc) : "NCCL error"));
    }
}

}  // namespace

Comm& comm()
{
    static Comm c;
    return c;
}

void comm_unique_id(void* id128)
{
    NcclId id;
    nccl_check(api().get_id(&id), "ncclGetUniqueId");
    memcpy(id128, id.internal, 128);
}

void comm_init(const void* id128, int rank, int nranks)
{
    Comm& c = comm();
    if (c.handle) comm_destroy();
    if (nranks < 1 || rank < 0 || rank >= nranks) throw ArgError("comm_init: bad rank / nranks");
    if (nranks == 1) { c.rank = 0; c.nranks = 1; return; }
    NcclId id;
    memcpy(id.internal, id128, 128);
    void* h = nullptr;
    nccl_check(api().init_rank(&h, nranks, id, rank), "ncclCommInitRank");
    c.handle = h; c.rank = rank; c.nranks = nranks;
}

void comm_destroy()
{
    peer_release();
    Comm& c = comm();
    if (c.handle) api().destroy(c.handle);
    c.handle = nullptr; c.rank = 0; c.nranks = 1;
}

void allreduce_sum(cudaStream_t s, float* buf, size_t count)
{
    Comm& c = comm();
    if (!c.active() || !count) return;
    nccl_check(api().allreduce(buf, buf, count, NCCL_F32, NCCL_SUM, c.handle, s), "ncclAllReduce");
}
void allreduce_sum(cudaStream_t s, double* buf, size_t count)
{
    Comm& c = comm();
    if (!c.active() || !count) return;
    nccl_check(api().allreduce(buf, buf, count, NCCL_F64, NCCL_SUM, c.handle, s), "ncclAllReduce");
}
double allreduce_sum_host(cudaStream_t s, double v)
{
    Comm& c = comm();
    if (!c.active()) return v;
    DevBuf<double> d(1);
    CUDA_CHECK(cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
    allreduce_sum(s, d.p, 1);
    CUDA_CHECK(cudaMemcpyAsync(&v, d.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return v;
}

void allgather_bytes(cudaStream_t s, const void* send, void* recv, size_t bytes)
{
    Comm& c = comm();
    if (!c.active()) { CUDA_CHECK(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s)); return; }
    if (!api().allgather) throw CodeError(B200ADMM_ENCCL, "NCCL library lacks ncclAllGather");
    nccl_check(api().allgather(send, recv, bytes, NCCL_U8, c.handle, s), "ncclAllGather");
}

// ---- peer-mapped exchange block ---------------------------------------------------------------------
static PeerBlock g_peer;

void peer_release()
{
    Comm& c = comm();
    for (int k = 0; k < 8; k++) {
        if (g_peer.peers[k] && k != c.rank) cudaIpcCloseMemHandle(g_peer.peers[k]);
        g_peer.peers[k] = nullptr;
    }
    if (g_peer.local) cudaFree(g_peer.local);
    cudaGetLastError();
    g_peer = PeerBlock();
}

PeerBlock& peer_block(cudaStream_t s, size_t floats)
{
    Comm& c = comm();
    if (!c.active() || c.nranks > 8) { g_peer.ok = false; return g_peer; }
    if (g_peer.ok && g_peer.floats >= floats) return g_peer;          // same decision on every rank: sizes are global
    peer_release();
    // plain cudaMalloc (not the block cache): the IPC handle names the whole allocation
    const size_t bytes = ((floats * sizeof(float) + ((size_t)2 << 20) - 1) >> 21) << 21;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (cudaMalloc((void**)&g_peer.local, bytes) != cudaSuccess) { cudaGetLastError(); g_peer.local = nullptr; ok = 0; }
    if (ok && cudaIpcGetMemHandle(&mine, g_peer.local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    // all ranks must walk through the same collectives whatever happened locally
    DevBuf<unsigned char> send(sizeof mine), recv(sizeof mine * (size_t)c.nranks);
    CUDA_CHECK(cudaMemcpyAsync(send.p, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
    allgather_bytes(s, send.p, recv.p, sizeof mine);
    std::vector<cudaIpcMemHandle_t> all(c.nranks);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), recv.p, sizeof mine * (size_t)c.nranks, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (allreduce_sum_host(s, (double)ok) < c.nranks - 0.5) ok = 0;    // somebody could not allocate / export
    if (ok) {
        for (int k = 0; k < c.nranks; k++) {
            if (k == c.rank) { g_peer.peers[k] = g_peer.local; continue; }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            g_peer.peers[k] = (float*)q;
        }
    }
    if (allreduce_sum_host(s, (double)ok) < c.nranks - 0.5) ok = 0;    // somebody could not map a peer
    if (!ok) { peer_release(); return g_peer; }
    g_peer.ok = true;
    g_peer.floats = bytes / sizeof(float);
    return g_peer;
}

}  // namespace b200
