// track2d_peer.cu -- the learner's only exchange between GPUs, over NVLink / NVSwitch peer memory instead of a library collective.
//
// Reference: main.py:102-116 + utils.py:36-44 (ensure_shared_grads): W Hogwild workers push their gradients into ONE shared model.
// Here every GPU holds a replica and applies the same update to the SUM of all ranks' gradients: one all-reduce of
// the flat fp32 gradient (~3.2 MB) per 20-step rollout.  As plain kernels it can live INSIDE the captured iteration graph, which the
// NCCL collective could not on this image (the capture hung), so a multi-GPU iteration is one graph replay with no host hop.
//
// Every rank owns a segment in its own HBM, allocated with cudaMalloc and exported through CUDA IPC:
//     data [n floats] | epoch | status | ready[8] | done[8]
// One all-reduce, on the caller's stream:
//   (1) wait until every peer has finished READING my previous data   (done[p] >= epoch: long true in practice)
//   (2) copy my gradient into my segment
//   (3) epoch += 1; tell every peer "my data of this epoch is ready" (release store, system scope, into THEIR ready[me]) and wait until
//       every peer has told me the same (acquire loads of MY ready[p])
//   (4) grad[i] = sum over ranks p = 0..W-1, in that order, of data_p[i]   (loads over NVLink; the same order on every rank, so the
//       replicas stay bit-identical)
//   (5) tell every peer "I have read your data" (done[me] in THEIR segment)
// The waits are bounded (about ten seconds of SM clock): a missing peer sets `status` instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/track2d.h"
#include "track2d_adam.cuh"

void t2d_set_error(const char *fmt, ...);
void t2d_count_launches(int n);

namespace {

constexpr int MAX_PEERS = 8;
constexpr long long SPIN_LIMIT = 20000000000ll;  // SM clocks (~10 s)

struct Flags {
    unsigned long long epoch;
    unsigned long long status;  // 0 ok; else 1 + the rank whose signal did not arrive
    unsigned long long ready[MAX_PEERS];
    unsigned long long done[MAX_PEERS];
};

struct PeerTable {
    const float *data[MAX_PEERS];
    Flags *flags[MAX_PEERS];
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_at_least(const unsigned long long *flag, unsigned long long v, Flags *mine, int who) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < v) {
        if (clock64() - t0 > SPIN_LIMIT) {
            mine->status = 1ull + (unsigned long long)who;
            break;
        }
        __nanosleep(100);
    }
}

// (1): lane p waits for peer p to have read my data of the current (= previous) epoch
__global__ void __launch_bounds__(32) peer_wait_done_kernel(PeerTable t) {
    Flags *mine = t.flags[t.rank];
    const int p = threadIdx.x;
    if (p < t.world && p != t.rank) wait_at_least(&mine->done[p], mine->epoch, mine, p);
}

// (3): the handshake that orders every rank's copy before every rank's loads
__global__ void __launch_bounds__(32) peer_ready_kernel(PeerTable t) {
    Flags *mine = t.flags[t.rank];
    __shared__ unsigned long long e;
    if (threadIdx.x == 0) {
        e = mine->epoch + 1ull;
        mine->epoch = e;
    }
    __syncwarp();
    const int p = threadIdx.x;
    if (p < t.world && p != t.rank) {
        __threadfence_system();  // my copy (earlier on this stream) is visible system-wide before the flag
        st_release_sys(&t.flags[p]->ready[t.rank], e);
        wait_at_least(&mine->ready[p], e, mine, p);
    }
}

// (4) + (5)
__global__ void __launch_bounds__(256) peer_sum_kernel(PeerTable t, float *__restrict__ grad, long long n, unsigned int *__restrict__ ticket) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < t.world; ++p) {
                const float4 v = __ldcv(reinterpret_cast<const float4 *>(t.data[p] + i));  // never from a stale L1 line
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            *reinterpret_cast<float4 *>(grad + i) = acc;
        } else {
            for (long long j = i; j < n; ++j) {
                float acc = 0.f;
                for (int p = 0; p < t.world; ++p) acc += __ldcv(t.data[p] + j);
                grad[j] = acc;
            }
        }
    }
    // the last block to finish tells the peers that this rank is done reading
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (last) *ticket = 0u;
    }
    __syncthreads();
    if (last && threadIdx.x < t.world && (int)threadIdx.x != t.rank) {
        const unsigned long long e = t.flags[t.rank]->epoch;
        st_release_sys(&t.flags[threadIdx.x]->done[t.rank], e);
    }
}

// (4) + (5) + the update: the exchange and SharedAdam.step in ONE kernel -- every thread adds up the ranks' gradients for its four
// parameters (same order everywhere), leaves the sum in `grad` (what an all-reduce would have left) and applies the update
__global__ void __launch_bounds__(256) peer_sum_adam_kernel(PeerTable t, float *__restrict__ grad, float *__restrict__ param, float *__restrict__ m,
                                                            float *__restrict__ v, float *__restrict__ vmax, long long n, float b1, float b2, float eps,
                                                            float grad_scale, const float *__restrict__ neg_step_dev, unsigned int *__restrict__ ticket) {
    const float neg_step = *neg_step_dev;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < t.world; ++p) {
                const float4 x = __ldcv(reinterpret_cast<const float4 *>(t.data[p] + i));
                G.x += x.x; G.y += x.y; G.z += x.z; G.w += x.w;
            }
            *reinterpret_cast<float4 *>(grad + i) = G;
            float4 P = *reinterpret_cast<float4 *>(param + i), M = *reinterpret_cast<float4 *>(m + i), V = *reinterpret_cast<float4 *>(v + i),
                   X = *reinterpret_cast<float4 *>(vmax + i);
            adam_one(P.x, G.x * grad_scale, M.x, V.x, X.x, b1, b2, eps, neg_step);
            adam_one(P.y, G.y * grad_scale, M.y, V.y, X.y, b1, b2, eps, neg_step);
            adam_one(P.z, G.z * grad_scale, M.z, V.z, X.z, b1, b2, eps, neg_step);
            adam_one(P.w, G.w * grad_scale, M.w, V.w, X.w, b1, b2, eps, neg_step);
            *reinterpret_cast<float4 *>(param + i) = P;
            *reinterpret_cast<float4 *>(m + i) = M;
            *reinterpret_cast<float4 *>(v + i) = V;
            *reinterpret_cast<float4 *>(vmax + i) = X;
        } else {
            for (long long j = i; j < n; ++j) {
                float g = 0.f;
                for (int p = 0; p < t.world; ++p) g += __ldcv(t.data[p] + j);
                grad[j] = g;
                adam_one(param[j], g * grad_scale, m[j], v[j], vmax[j], b1, b2, eps, neg_step);
            }
        }
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (last) *ticket = 0u;
    }
    __syncthreads();
    if (last && threadIdx.x < t.world && (int)threadIdx.x != t.rank) {
        const unsigned long long e = t.flags[t.rank]->epoch;
        st_release_sys(&t.flags[threadIdx.x]->done[t.rank], e);
    }
}

}  // namespace

struct track2d_peer {
    int rank, world, device, connected;
    long long n;
    size_t bytes;
    uint8_t *seg;         // my segment (cudaMalloc)
    void *opened[MAX_PEERS];  // cudaIpcOpenMemHandle results (to close), null for local pointers
    unsigned int *ticket;
    PeerTable table;
};

namespace {
struct Guard {
    int prev;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};
size_t data_bytes(long long n) { return ((size_t)n * 4 + 255) / 256 * 256; }
int cuda_fail(const char *what, cudaError_t e) {
    t2d_set_error("%s: %s", what, cudaGetErrorString(e));
    return T2D_E_CUDA;
}
void set_entry(track2d_peer *p, int r, uint8_t *base) {
    p->table.data[r] = reinterpret_cast<const float *>(base);
    p->table.flags[r] = reinterpret_cast<Flags *>(base + data_bytes(p->n));
}
}  // namespace

extern "C" int track2d_peer_create(int32_t rank, int32_t world, int64_t n_floats, int32_t device, track2d_peer **out) {
    if (!out || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || n_floats < 1) {
        t2d_set_error("track2d_peer_create: bad argument (1 <= world <= %d)", MAX_PEERS);
        return T2D_E_INVALID;
    }
    Guard g(device);
    track2d_peer *p = (track2d_peer *)calloc(1, sizeof(track2d_peer));
    if (!p) return T2D_E_INVALID;
    p->rank = rank; p->world = world; p->device = device; p->n = n_floats;
    p->bytes = data_bytes(n_floats) + sizeof(Flags);
    cudaError_t e = cudaMalloc(&p->seg, p->bytes);
    if (e == cudaSuccess) e = cudaMemset(p->seg, 0, p->bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p->ticket, 0, sizeof(unsigned int));
    if (e != cudaSuccess) {
        if (p->seg) cudaFree(p->seg);
        free(p);
        return cuda_fail("track2d_peer_create", e);
    }
    {   // load the kernels now: with lazy module loading the FIRST launch of a function can synchronise with running kernels, and these
        // kernels wait for each other across streams / processes
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, peer_wait_done_kernel);
        cudaFuncGetAttributes(&fa, peer_ready_kernel);
        cudaFuncGetAttributes(&fa, peer_sum_kernel);
        cudaFuncGetAttributes(&fa, peer_sum_adam_kernel);
        t2d_preload_optim_kernels();
    }
    p->table.rank = rank; p->table.world = world;
    set_entry(p, rank, p->seg);
    p->connected = world == 1;
    *out = p;
    return T2D_OK;
}

extern "C" int track2d_peer_handle(track2d_peer *p, uint8_t *handle64_out) {
    if (!p || !handle64_out) { t2d_set_error("track2d_peer_handle: bad argument"); return T2D_E_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    Guard g(p->device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p->seg);
    if (e != cudaSuccess) return cuda_fail("track2d_peer_handle", e);
    memcpy(handle64_out, &h, 64);
    return T2D_OK;
}

extern "C" int track2d_peer_segment(track2d_peer *p, void **segment_out) {
    if (!p || !segment_out) { t2d_set_error("track2d_peer_segment: bad argument"); return T2D_E_INVALID; }
    *segment_out = p->seg;
    return T2D_OK;
}

extern "C" int track2d_peer_connect(track2d_peer *p, const uint8_t *handles_world_x_64) {
    if (!p || !handles_world_x_64) { t2d_set_error("track2d_peer_connect: bad argument"); return T2D_E_INVALID; }
    Guard g(p->device);
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles_world_x_64 + 64 * r, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail("track2d_peer_connect (cudaIpcOpenMemHandle)", e);
        p->opened[r] = ptr;
        set_entry(p, r, (uint8_t *)ptr);
    }
    p->connected = 1;
    return T2D_OK;
}

// same-process peers (tests; several GPUs driven by one process): raw segment pointers instead of IPC handles
extern "C" int track2d_peer_connect_local(track2d_peer *p, void *const *segments_world) {
    if (!p || !segments_world) { t2d_set_error("track2d_peer_connect_local: bad argument"); return T2D_E_INVALID; }
    for (int r = 0; r < p->world; ++r)
        if (r != p->rank) set_entry(p, r, (uint8_t *)segments_world[r]);
    p->connected = 1;
    return T2D_OK;
}

extern "C" int track2d_peer_allreduce(track2d_peer *p, float *grad_dev, void *stream) {
    if (!p || !grad_dev || ((uintptr_t)grad_dev & 15u)) { t2d_set_error("track2d_peer_allreduce: bad argument (16-byte aligned gradient)"); return T2D_E_INVALID; }
    if (!p->connected) { t2d_set_error("track2d_peer_allreduce: call track2d_peer_connect first"); return T2D_E_STATE; }
    if (p->world == 1) return T2D_OK;
    Guard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    peer_wait_done_kernel<<<1, 32, 0, st>>>(p->table);
    cudaError_t e = cudaMemcpyAsync(p->seg, grad_dev, (size_t)p->n * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail("track2d_peer_allreduce", e);
    peer_ready_kernel<<<1, 32, 0, st>>>(p->table);
    long long blocks = (p->n / 4 + 255) / 256;
    if (blocks > 296) blocks = 296;
    if (blocks < 1) blocks = 1;
    peer_sum_kernel<<<(int)blocks, 256, 0, st>>>(p->table, grad_dev, p->n, p->ticket);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail("track2d_peer_allreduce", e);
    t2d_count_launches(3);
    return T2D_OK;
}

// The exchange fused with the update it feeds: same result, bit for bit, as track2d_peer_allreduce followed by track2d_sharedadam_step
// (without clipping -- a global gradient norm needs the whole sum first), in one launch less and without re-reading the summed gradient.
extern "C" int track2d_peer_sharedadam_step(track2d_peer *p, float *param, float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq,
                                            double lr, double beta1, double beta2, double eps, double grad_scale, float *norm_scratch,
                                            int64_t *step_dev, void *stream) {
    if (!p || !param || !grad || !exp_avg || !exp_avg_sq || !max_exp_avg_sq || !norm_scratch || !step_dev ||
        ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)max_exp_avg_sq) & 15u) != 0)) {
        t2d_set_error("track2d_peer_sharedadam_step: bad argument (16-byte aligned buffers, device-resident step counter)");
        return T2D_E_INVALID;
    }
    if (!p->connected) { t2d_set_error("track2d_peer_sharedadam_step: call track2d_peer_connect first"); return T2D_E_STATE; }
    Guard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (p->world > 1) {
        peer_wait_done_kernel<<<1, 32, 0, st>>>(p->table);
        e = cudaMemcpyAsync(p->seg, grad, (size_t)p->n * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_fail("track2d_peer_sharedadam_step", e);
        peer_ready_kernel<<<1, 32, 0, st>>>(p->table);
    } else {
        e = cudaMemcpyAsync(p->seg, grad, (size_t)p->n * 4, cudaMemcpyDeviceToDevice, st);  // world 1: the "sum" of one copy
        if (e != cudaSuccess) return cuda_fail("track2d_peer_sharedadam_step", e);
    }
    e = t2d_launch_adam_prep(reinterpret_cast<long long *>(step_dev), norm_scratch + 1, lr, beta1, beta2, st);
    if (e != cudaSuccess) return cuda_fail("track2d_peer_sharedadam_step", e);
    long long blocks = (p->n / 4 + 255) / 256;
    if (blocks > 296) blocks = 296;
    if (blocks < 1) blocks = 1;
    peer_sum_adam_kernel<<<(int)blocks, 256, 0, st>>>(p->table, grad, param, exp_avg, exp_avg_sq, max_exp_avg_sq, p->n, (float)beta1, (float)beta2,
                                                      (float)eps, (float)grad_scale, norm_scratch + 1, p->ticket);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail("track2d_peer_sharedadam_step", e);
    t2d_count_launches(p->world > 1 ? 4 : 2);
    return T2D_OK;
}

extern "C" int track2d_peer_status(track2d_peer *p, uint64_t *status_out) {
    if (!p || !status_out) { t2d_set_error("track2d_peer_status: bad argument"); return T2D_E_INVALID; }
    Guard g(p->device);
    Flags f;
    cudaError_t e = cudaMemcpy(&f, p->seg + data_bytes(p->n), sizeof(Flags), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail("track2d_peer_status", e);
    *status_out = f.status;
    return T2D_OK;
}

extern "C" void track2d_peer_destroy(track2d_peer *p) {
    if (!p) return;
    Guard g(p->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < MAX_PEERS; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
    if (p->seg) cudaFree(p->seg);
    if (p->ticket) cudaFree(p->ticket);
    free(p);
}
