// Micro-benchmarks of the primitives the persistent decode kernel is built from (run on a B200 via gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I blama_b200/csrc tools/micro/lat.cu -o build/lat && build/lat
#include <cstdio>
#include <vector>
#include "mega_decode.cuh"
using namespace blk;

// the grid barrier the first version of the persistent kernel used (kept here for the measurement)
__device__ __forceinline__ void mg_grid_arrive(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) red_release_add(bar, 1u);
}
__device__ __forceinline__ void mg_grid_wait(const unsigned int* bar, unsigned int target) {
    if (threadIdx.x == 0) {
        while (ld_relaxed_u32(bar) < target) { }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

__global__ void k_quant(float* src, int8_t* out, long long* t, int iters) {
    __shared__ __align__(16) int8_t sq[4096]; __shared__ float sd[64]; __shared__ int16_t sbs[512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MgV8 v = mg_load8(src + warp * 256 + lane * 8);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { mg_quantize_regs(v.a, v.b, ACT_Q8_K, warp, sq, sd, sbs); v.a.x += 1.0f; }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { t[0] = t1 - t0; out[0] = sq[5]; }
}
// chain of dependent ALU ops / shuffles / double ops
__global__ void k_chain(float* io, long long* t, int iters) {
    float x = io[threadIdx.x]; double d = x; int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = x * 1.0001f + 0.5f;
    long long t1 = clock64();
    for (int i = 0; i < iters; i++) x += __shfl_xor_sync(0xffffffffu, x, 1);
    long long t2 = clock64();
    for (int i = 0; i < iters; i++) d = d * 1.0001 + 0.5;
    long long t3 = clock64();
    for (int i = 0; i < iters; i++) d = 1.0 / (d + 3.0);
    long long t4 = clock64();
    for (int i = 0; i < iters; i++) x = __fdiv_rn(1.0f, x + 3.0f);
    long long t5 = clock64();
    for (int i = 0; i < iters; i++) x = expf(x * 0.001f);
    long long t6 = clock64();
    io[threadIdx.x] = x + (float)d;
    if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; }
    (void)lane;
}
// grid barrier round trip + read of data written by the other CTAs, 148 CTAs x 512 threads, cooperative
__global__ void k_barrier(unsigned int* bar, float* data, long long* t, int iters, int n_cta) {
    unsigned int target = 0;
    long long acc_bar = 0, acc_ld = 0;
    float s = 0.f;
    for (int i = 0; i < iters; i++) {
        data[(size_t)(i & 1) * 65536 + blockIdx.x * 512 + threadIdx.x] = (float)i + s * 1e-9f;
        long long t0 = clock64();
        mg_grid_arrive(bar); target += n_cta; mg_grid_wait(bar, target);
        long long t1 = clock64();
        const int other = (blockIdx.x + 37) % n_cta;
        s += __ldcg(data + (size_t)(i & 1) * 65536 + other * 512 + threadIdx.x);
        long long t2 = clock64() + (long long)(s == 12345.f);
        acc_bar += t1 - t0; acc_ld += t2 - t1;
    }
    if (threadIdx.x == 0) { t[blockIdx.x * 2] = acc_bar / iters; t[blockIdx.x * 2 + 1] = acc_ld / iters; }
    if (s == -1.f) data[0] = s;
}
// barrier variants: V=0 as shipped; 1: no acquire fence on the wait side; 2: no fences at all (floor; not a correct barrier);
// 3: LL-style exchange, no barrier: every value travels with an epoch flag in one 8-byte store, readers poll the data itself
template <int V>
__global__ void k_barrier_v(unsigned int* bar, float* data, long long* t, int iters, int n_cta) {
    unsigned int target = 0;
    long long acc = 0;
    float s = 0.f;
    uint2* d2 = reinterpret_cast<uint2*>(data);
    for (int i = 0; i < iters; i++) {
        const int other = (blockIdx.x + 37) % n_cta;
        long long t0 = clock64();
        if (V == 3) {
            d2[blockIdx.x * 512 + threadIdx.x] = make_uint2(__float_as_uint((float)i + s * 1e-9f), (unsigned)i + 1u);
            uint2 v;
            int spins = 0;
            do { asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(d2 + other * 512 + threadIdx.x) : "memory"); } while (!__all_sync(0xffffffffu, v.y >= (unsigned)i + 1u) && ++spins < 100000);
            s += __uint_as_float(v.x);
            __syncthreads();
        } else {
            data[(size_t)(i & 1) * 65536 + blockIdx.x * 512 + threadIdx.x] = (float)i + s * 1e-9f;
            __syncthreads();
            if (threadIdx.x == 0) {
                if (V == 2) atomicAdd(bar, 1u); else red_release_add(bar, 1u);
                target += n_cta;
                int spins = 0;
                while (ld_relaxed_u32(bar) < target && ++spins < 100000) { }
                if (V == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncthreads();
            s += __ldcg(data + (size_t)(i & 1) * 65536 + other * 512 + threadIdx.x);
        }
        long long t2 = clock64() + (long long)(s == 12345.f);
        acc += t2 - t0;
    }
    if (threadIdx.x == 0) t[blockIdx.x] = acc / iters;
    if (s == -1.f) data[0] = s;
}
template <int V> void run_v(unsigned int* bar, float* data, long long* t, const char* name) {
    int n_cta = 148, iters = 200; long long h[148];
    void* args[] = {&bar, &data, &t, &iters, &n_cta};
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(bar, 0, 64); cudaMemset(data, 0, 4 << 20);
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k_barrier_v<V>, dim3(n_cta), dim3(512), args, 0, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, t, 148 * 8, cudaMemcpyDeviceToHost);
        double b = 0; for (int i = 0; i < 148; i++) b += h[i];
        printf("%-60s %.0f cycles (store + sync + read of another CTA's data) %s\n", name, b / 148, cudaGetErrorString(e));
    }
}
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    float* src; int8_t* out; long long* t; unsigned int* bar; float* data;
    cudaMalloc(&src, 1 << 20); cudaMalloc(&out, 64); cudaMalloc(&t, 4096 * 8); cudaMalloc(&bar, 64); cudaMalloc(&data, 4 << 20);
    cudaMemset(src, 0x3c, 1 << 20); cudaMemset(bar, 0, 64); cudaMemset(data, 0, 4 << 20);
    long long h[512];
    for (int rep = 0; rep < 2; rep++) {
        k_quant<<<1, 512>>>(src, out, t, 100); cudaMemcpy(h, t, 8, cudaMemcpyDeviceToHost);
        printf("quantize_regs (16 warps, hot loop): %.1f cycles per call\n", h[0] / 100.0);
        k_quant<<<1, 32>>>(src, out, t, 100); cudaMemcpy(h, t, 8, cudaMemcpyDeviceToHost);
        printf("quantize_regs (1 warp, hot loop):   %.1f cycles per call\n", h[0] / 100.0);
        k_quant<<<1, 512>>>(src, out, t, 1); cudaMemcpy(h, t, 8, cudaMemcpyDeviceToHost);
        printf("quantize_regs (16 warps, single call): %lld cycles\n", h[0]);
    }
    k_chain<<<1, 512>>>(src, t, 1000); cudaMemcpy(h, t, 48, cudaMemcpyDeviceToHost);
    printf("dependent chain, cycles per op (16 warps): ffma %.1f  shfl+fadd %.1f  dfma %.1f  ddiv %.1f  fdiv_rn %.1f  expf %.1f\n", h[0] / 1e3, h[1] / 1e3, h[2] / 1e3, h[3] / 1e3, h[4] / 1e3, h[5] / 1e3);
    k_chain<<<1, 32>>>(src, t, 1000); cudaMemcpy(h, t, 48, cudaMemcpyDeviceToHost);
    printf("dependent chain, cycles per op (1 warp):   ffma %.1f  shfl+fadd %.1f  dfma %.1f  ddiv %.1f  fdiv_rn %.1f  expf %.1f\n", h[0] / 1e3, h[1] / 1e3, h[2] / 1e3, h[3] / 1e3, h[4] / 1e3, h[5] / 1e3);
    int n_cta = 148, iters = 200;
    void* args[] = {&bar, &data, &t, &iters, &n_cta};
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(bar, 0, 64);
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k_barrier, dim3(n_cta), dim3(512), args, 0, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, t, 148 * 16, cudaMemcpyDeviceToHost);
        double b = 0, l = 0; for (int i = 0; i < 148; i++) { b += h[2 * i]; l += h[2 * i + 1]; }
        printf("grid barrier (148 x 512, idle memory system): %.0f cycles; ldcg of another CTA's fresh data: %.0f cycles (%s)\n", b / 148, l / 148, cudaGetErrorString(e));
    }
    run_v<0>(bar, data, t, "barrier: red.release + relaxed poll + acq_rel fence");
    run_v<1>(bar, data, t, "barrier: red.release + relaxed poll, no acquire fence");
    run_v<2>(bar, data, t, "barrier: plain atomic, no fences (floor, not correct)");
    run_v<3>(bar, data, t, "LL exchange: {value, epoch} 8-byte words, readers poll the data");
    return 0;
}
