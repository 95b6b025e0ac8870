// od_exchange.cu — peer-memory plumbing for the one cross-GPU exchange of the path: the 8 fp64
// loss partial sums of a step.  Instead of an NCCL all-reduce node between the loss kernel and the finalize
// kernel, the LAST CTA of k_pos_loss_tiles (od_loss.cu) pushes this GPU's sums into every peer's exchange
// region with plain stores over NVLink (P2P), waits for the peers' pushes, adds the W contributions in rank
// order (same bits on every rank) and finalizes the losses — compute and collective in one kernel.
//
// This file only allocates the regions and passes CUDA IPC handles around (one process per GPU):
//   sihl_od_exchange_create  : cudaMalloc + zero a block of n_regions regions, export its IPC handle
//   sihl_od_exchange_open    : map a peer's block (enables peer access lazily)
//   sihl_od_exchange_close / _destroy
// The handles travel between the processes through torch.distributed (sihl_b200/dist.py); nothing here is on
// the per-step path.
//
// Region layout (8-byte words; W = world size; parity = step & 1 so a rank that is one step ahead never
// overwrites what a slower rank still reads — a rank cannot be two steps ahead, it needs every peer's push):
//   [0, 16 W)            sums   [parity][source rank][8]   double
//   [16 W, 18 W)         flags  [parity][source rank]      u64: the step number whose sums are complete
//   [18 W]               this GPU's own step counter       u64
//   [18 W + 1]           wait limit in ns (0 = 120 s default)   u64   sihl_od_exchange_set_timeout
//   [18 W + 2]           sticky error: first step whose wait timed out (0 = none)   sihl_od_exchange_status
#include "od_common.cuh"

using namespace sihl;

extern "C" size_t sihl_od_exchange_region_bytes(int world)
{
    if (world < 1 || world > SIHL_OD_MAX_PEERS) return 0;
    const size_t words = (size_t)18 * world + 3;
    return (words * 8 + 255) / 256 * 256;
}

extern "C" int sihl_od_exchange_create(int world, int n_regions, void **block, unsigned char *ipc_handle_out)
{
    SIHL_CHECK_ARG(world >= 1 && world <= SIHL_OD_MAX_PEERS, "world=%d outside 1..%d", world, SIHL_OD_MAX_PEERS);
    SIHL_CHECK_ARG(n_regions >= 1 && block && ipc_handle_out, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == SIHL_OD_IPC_HANDLE_BYTES, "handle size");
    const size_t bytes = sihl_od_exchange_region_bytes(world) * (size_t)n_regions;
    void *ptr = nullptr;
    int rc = cuda_status(cudaMalloc(&ptr, bytes), "cudaMalloc(exchange block)");
    if (rc) return rc;
    rc = cuda_status(cudaMemset(ptr, 0, bytes), "cudaMemset(exchange block)");
    if (!rc) rc = cuda_status(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    cudaIpcMemHandle_t h;
    if (!rc) rc = cuda_status(cudaIpcGetMemHandle(&h, ptr), "cudaIpcGetMemHandle");
    if (rc) { cudaFree(ptr); return rc; }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *block = ptr;
    return SIHL_OD_OK;
}

extern "C" int sihl_od_exchange_open(const unsigned char *ipc_handle, void **peer_block)
{
    SIHL_CHECK_ARG(ipc_handle && peer_block, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    return cuda_status(cudaIpcOpenMemHandle(peer_block, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

extern "C" int sihl_od_exchange_close(void *peer_block)
{
    if (peer_block == nullptr) return SIHL_OD_OK;
    return cuda_status(cudaIpcCloseMemHandle(peer_block), "cudaIpcCloseMemHandle");
}

extern "C" int sihl_od_exchange_destroy(void *block)
{
    if (block == nullptr) return SIHL_OD_OK;
    return cuda_status(cudaFree(block), "cudaFree(exchange block)");
}

// Device-side barrier over the node's GPUs on one exchange region (one CTA, thread q talks to rank q): every rank
// publishes its next step number into every peer's flag word with a release store and waits, with acquire loads, until
// every peer's flag has arrived.  Enqueued in front of a timed region it makes all GPUs start within an NVLink
// round trip of each other WITHOUT a host barrier (whose rank skew would land inside the region); same region layout,
// timeout and sticky error word as the fused loss-sum exchange (od_loss.cu), so a region must be used for one or the
// other, never both.
__global__ void __launch_bounds__(32) k_exchange_barrier(unsigned long long *const *peer, int W, int rank)
{
    __shared__ unsigned long long s_step;
    const int tid = threadIdx.x;
    unsigned long long *mine = peer[rank];
    if (tid == 0) s_step = ++mine[18 * W];
    __syncwarp();
    const unsigned long long step = s_step;
    const int parity = (int)(step & 1ull);
    if (tid < W) {
        unsigned long long *theirs = peer[tid];
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs + 16 * W + parity * W + rank), "l"(step) : "memory");
        const unsigned long long *flag = mine + 16 * W + parity * W + tid;
        unsigned long long t0, now, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned long long limit = *reinterpret_cast<volatile unsigned long long *>(mine + 18 * W + 1);
        if (limit == 0ull) limit = 120000000000ull;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v == step) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > limit) { atomicCAS(mine + 18 * W + 2, 0ull, step); break; }
            __nanosleep(64);
        }
    }
}

extern "C" int sihl_od_exchange_barrier(void *const *peer_regions, int world, int rank, void *stream)
{
    SIHL_CHECK_ARG(peer_regions && world >= 1 && world <= SIHL_OD_MAX_PEERS && rank >= 0 && rank < world, "bad arguments");
    if (world == 1) return SIHL_OD_OK;
    k_exchange_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long *const *>(peer_regions), world, rank);
    SIHL_CHECK_LAUNCH("k_exchange_barrier");
    return SIHL_OD_OK;
}

extern "C" int sihl_od_exchange_set_timeout(void *region, int world, uint64_t timeout_ns)
{
    SIHL_CHECK_ARG(region && world >= 1 && world <= SIHL_OD_MAX_PEERS, "bad arguments");
    return cuda_status(cudaMemcpy(reinterpret_cast<unsigned long long *>(region) + 18 * world + 1, &timeout_ns, sizeof(timeout_ns),
                                  cudaMemcpyHostToDevice), "cudaMemcpy(exchange timeout)");
}

extern "C" int sihl_od_exchange_status(const void *region, int world, uint64_t *timed_out_step, uint64_t *steps_done)
{
    SIHL_CHECK_ARG(region && world >= 1 && world <= SIHL_OD_MAX_PEERS && timed_out_step, "bad arguments");
    unsigned long long w[3] = {0, 0, 0};
    int rc = cuda_status(cudaMemcpy(w, reinterpret_cast<const unsigned long long *>(region) + 18 * world, sizeof(w),
                                    cudaMemcpyDeviceToHost), "cudaMemcpy(exchange status)");
    if (rc) return rc;
    *timed_out_step = w[2];
    if (steps_done) *steps_done = w[0];
    return SIHL_OD_OK;
}
