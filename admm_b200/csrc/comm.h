// comm.h -- the one collective this library needs: sum all-reduce over NVLink via NCCL.
//
// One process per GPU; the host program (bench.py / the R session's launcher) creates the
// communicator with b200admm_comm_id + b200admm_comm_init.  NCCL is bound at run time with
// dlopen("libnccl.so.2") so that libb200admm.so has no link-time dependency on it and reuses
// the copy a host process may already have loaded.  With no communicator installed every
// call below is a no-op on a "world" of one rank.
#pragma once
#include "common.cuh"

namespace b200 {

struct Comm {
    int rank = 0;
    int nranks = 1;
    void* handle = nullptr;       // ncclComm_t
    bool suspended = false;       // b200admm_comm_suspend: calls behave as on a single GPU while set
    bool active() const { return handle != nullptr && nranks > 1 && !suspended; }
};
Comm& comm();

void comm_unique_id(void* id128);
void comm_init(const void* id128, int rank, int nranks);
void comm_destroy();

void allreduce_sum(cudaStream_t s, float* buf, size_t count);
void allreduce_sum(cudaStream_t s, double* buf, size_t count);
// host scalars through a small device staging buffer (setup only)
double allreduce_sum_host(cudaStream_t s, double v);
// recv (nranks * bytes) <- every rank's send (bytes); device buffers
void allgather_bytes(cudaStream_t s, const void* send, void* recv, size_t bytes);

// One device block per rank that every other rank can store into (cudaIpc mapping over NVLink): the exchange
// buffer of the sharded iteration kernel (fadmm_tall.cu).  peer_block is COLLECTIVE: all ranks call it with the
// same size.  The block and its mappings are kept until the size grows or the communicator goes away.
struct PeerBlock {
    bool ok = false;
    float* local = nullptr;
    size_t floats = 0;
    float* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
PeerBlock& peer_block(cudaStream_t s, size_t floats);
void peer_release();

}  // namespace b200
