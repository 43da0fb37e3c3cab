// Host side of the peer-memory exchange (include/simgan_b200.h: sg_dp_*): allocation, CUDA IPC handle
// exchange (the handles travel between processes through the caller, e.g. torch.distributed.all_gather_object).
#include <string.h>

#include "sg_dp.cuh"

namespace sg {
struct DpCtx {
    int rank, world, cap;
    void* local;                           // cudaMalloc'ed exchange buffer of this rank
    void* peers[kDpMaxWorld];              // mapped peer buffers (peers[rank] == local)
    unsigned int* error_flag;
    bool opened;
};

DpView dp_view(const void* ctx) {
    const DpCtx* c = (const DpCtx*)ctx;
    DpView v;
    memset(&v, 0, sizeof(v));
    v.rank = c->rank; v.world = c->world; v.cap = c->cap;
    for (int r = 0; r < c->world; ++r) {
        v.gather[r] = (float*)c->peers[r];
        v.flags[r] = (unsigned int*)((char*)c->peers[r] + dp_gather_floats(c->world, c->cap) * sizeof(float));
    }
    v.error_flag = c->error_flag;
    return v;
}
}  // namespace sg

using namespace sg;

extern "C" {
#pragma GCC visibility push(default)

int sg_dp_create(int rank, int world, int max_floats, void** out_ctx) {
    SG_REQUIRE(out_ctx && world >= 2 && world <= kDpMaxWorld && rank >= 0 && rank < world && max_floats > 0,
               "sg_dp_create: bad arguments (world must be 2..%d)", kDpMaxWorld);
    DpCtx* c = new DpCtx();
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world; c->cap = round_up(max_floats, 4);
    const size_t bytes = dp_bytes(world, c->cap);
    SG_CUDA(cudaMalloc(&c->local, bytes));
    SG_CUDA(cudaMemset(c->local, 0, bytes));
    SG_CUDA(cudaDeviceSynchronize());
    c->peers[rank] = c->local;
    c->error_flag = (unsigned int*)((char*)c->local + bytes - 64);
    *out_ctx = c;
    return SG_OK;
}

int sg_dp_local_handle(void* ctx, unsigned char* out64) {
    SG_REQUIRE(ctx && out64, "sg_dp_local_handle: null pointer");
    DpCtx* c = (DpCtx*)ctx;
    cudaIpcMemHandle_t h;
    SG_CUDA(cudaIpcGetMemHandle(&h, c->local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &h, 64);
    return SG_OK;
}

int sg_dp_open_peers(void* ctx, const unsigned char* handles) {
    SG_REQUIRE(ctx && handles, "sg_dp_open_peers: null pointer");
    DpCtx* c = (DpCtx*)ctx;
    SG_REQUIRE(!c->opened, "sg_dp_open_peers: already opened");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        SG_CUDA(cudaIpcOpenMemHandle(&c->peers[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->opened = true;
    return SG_OK;
}

int sg_dp_capacity(void* ctx) { return ctx ? ((DpCtx*)ctx)->cap : -1; }

int sg_dp_destroy(void* ctx) {
    if (!ctx) return SG_OK;
    DpCtx* c = (DpCtx*)ctx;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
    if (c->local) cudaFree(c->local);
    cudaGetLastError();
    delete c;
    return SG_OK;
}

#pragma GCC visibility pop
}  // extern "C"
