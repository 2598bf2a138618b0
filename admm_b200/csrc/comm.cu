// comm.cu -- NCCL binding (run-time dlopen) for the row-sharded multi-GPU paths.
#include "../../include/b200admm.h"
#include "comm.h"
#include <dlfcn.h>
#include <cstring>

namespace b200 {

namespace {

// the handful of NCCL entry points we use, with the ABI of nccl.h 2.27 / 2.28
typedef struct { char internal[128]; } NcclId;
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
constexpr int NCCL_F32 = 7, NCCL_F64 = 8, NCCL_SUM = 0;

struct NcclApi {
    void* lib = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
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

}  // namespace b200
