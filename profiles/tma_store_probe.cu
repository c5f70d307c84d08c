// What does the shared->global TMA bulk-store path sustain on this B200, as a function of the store size, the
// stores in flight per CTA and the CTAs per SM?  (k_obs is exactly this loop plus row composition.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_store_probe profiles/tma_store_probe.cu && /tmp/tma_store_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void bulk_store(void *g, const void *s, uint32_t bytes) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g), "r"(sa), "r"(bytes) : "memory");
}

template <int INFLIGHT>
__global__ void k_tma(char *out, size_t total, uint32_t chunk, int linear) {
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t n_chunks = total / chunk;
    for (uint32_t i = threadIdx.x; i < chunk * INFLIGHT / 4; i += blockDim.x) ((uint32_t *)smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    // linear: chunk index = k * grid + block (neighbouring CTAs write neighbouring chunks); else each CTA owns a contiguous range
    const size_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
    int b = 0;
    for (size_t k = 0; k < per; k++) {
        const size_t c = linear ? k * gridDim.x + blockIdx.x : blockIdx.x * per + k;
        if (c >= n_chunks) break;
        asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(INFLIGHT - 1) : "memory");
        bulk_store(out + c * chunk, smem + (size_t)b * chunk, chunk);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        b = (b + 1) % INFLIGHT;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// dynamic distribution: a CTA takes the next block of `per_grab` chunks from a global ticket counter
template <int INFLIGHT>
__global__ void k_tma_dyn(char *out, size_t total, uint32_t chunk, int per_grab, unsigned long long *ticket) {
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t n_chunks = total / chunk;
    for (uint32_t i = threadIdx.x; i < chunk * INFLIGHT / 4; i += blockDim.x) ((uint32_t *)smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    int b = 0;
    unsigned long long next = atomicAdd(ticket, 1ull);
    while (true) {
        const size_t first = (size_t)next * per_grab;
        if (first >= n_chunks) break;
        next = atomicAdd(ticket, 1ull);                      // prefetch the next ticket under this block's stores
        for (int k = 0; k < per_grab && first + k < n_chunks; k++) {
            asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(INFLIGHT - 1) : "memory");
            bulk_store(out + (first + k) * chunk, smem + (size_t)b * chunk, chunk);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            b = (b + 1) % INFLIGHT;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void k_stg(uint4 *out, size_t n16) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

template <int INFLIGHT>
static void run_tma(char *d, size_t total, uint32_t chunk, int ctas_per_sm, int linear, int sms) {
    const size_t smem = (size_t)chunk * INFLIGHT;
    if (smem * ctas_per_sm > 220 * 1024) return;
    cudaFuncSetAttribute(k_tma<INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_tma<INFLIGHT><<<sms * ctas_per_sm, 128, smem>>>(d, total, chunk, linear);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const size_t written = total / chunk * chunk;
    printf("tma  chunk %6u B  inflight %d  ctas/sm %d  %-10s %.3f ms  %.0f GB/s  (%s)\n", chunk, INFLIGHT, ctas_per_sm,
           linear ? "interleaved" : "ranges", best, written / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t total = 2552000000ull / 16 * 16;
    char *d; cudaMalloc(&d, total + 1024);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {sms * 8, sms * 32, sms * 128}) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0); k_stg<<<blocks, 256>>>((uint4 *)d, total / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        printf("stg.128 grid %6d x 256: %.3f ms  %.0f GB/s\n", blocks, best, total / best / 1e6);
    }
    unsigned long long *ticket; cudaMalloc(&ticket, 8);
    for (int per_grab : {1, 8, 32})
        for (int cps : {1, 2}) {
            const uint32_t chunk = 37856; const size_t smem = (size_t)chunk * 2;
            cudaFuncSetAttribute(k_tma_dyn<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            float best = 1e9f;
            for (int rep = 0; rep < 5; rep++) {
                cudaMemset(ticket, 0, 8);
                cudaEventRecord(e0);
                k_tma_dyn<2><<<sms * cps, 128, smem>>>(d, total, chunk, per_grab, ticket);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
            }
            printf("tma dynamic  chunk %u B x %2d per ticket  inflight 2  ctas/sm %d  %.3f ms  %.0f GB/s (%s)\n", chunk, per_grab,
                   cps, best, total / chunk * chunk / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (int linear : {1})
        for (uint32_t chunk : {37856u}) {
            for (int cps : {1, 2, 4, 8}) {
                run_tma<1>(d, total, chunk, cps, linear, sms);
                run_tma<2>(d, total, chunk, cps, linear, sms);
                run_tma<4>(d, total, chunk, cps, linear, sms);
            }
        }
    return 0;
}
