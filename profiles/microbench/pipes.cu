// Micro-benchmarks that fix the roofline denominators for the NPHD scan on B200 (sm_100a):
//   * issue rate of POPC / LOP3 / IADD3 per SM per clock (register-only loops)
//   * whether POPC overlaps with LOP3/IADD3 (separate pipes) - decides if carry-save
//     compression (fewer POPC, more LOP3) can pay
//   * HBM read-only streaming bandwidth with 128-bit loads over a buffer >> L2
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// mode 0: 8 POPC / iter          mode 1: 8 LOP3 / iter       mode 2: 8 IADD3 / iter
// mode 3: 8 POPC + 8 LOP3        mode 4: 8 POPC + 16 LOP3    mode 5: 8 POPC + 8 LOP3 + 8 IADD
// mode 6: 8 POPC + 24 LOP3       mode 7: 8 IMAD              mode 8: 8 POPC + 8 LOP3 + 8 IMAD
template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a[8], p[8], l[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; p[i] = a[i] * 3; l[i] = a[i] ^ 0x5555u; s[i] = i; }
    uint32_t k1 = seed | 1, k2 = seed ^ 0xdeadbeefu;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 8)
                asm volatile("popc.b32 %0, %0;" : "+r"(p[i]));
            if (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 8)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(k1), "r"(k2));
            if (MODE == 4 || MODE == 6)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(s[i]) : "r"(k1), "r"(k2));
            if (MODE == 6)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k2), "r"(k1));
            if (MODE == 2 || MODE == 5)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(k1));
            if (MODE == 9 && (i & 1))   // 4 POPC + 16 LOP3 per iter (popc:lop3 = 1:4)
                asm volatile("popc.b32 %0, %0;" : "+r"(p[i]));
            if (MODE == 9) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(k1), "r"(k2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(s[i]) : "r"(k1), "r"(k2));
            }
            if (MODE == 10) {           // 8 POPC + 4 LOP3 (popc:lop3 = 2:1)
                asm volatile("popc.b32 %0, %0;" : "+r"(p[i]));
                if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(k1), "r"(k2));
            }
            if (MODE == 7 || MODE == 8)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(k1), "r"(k2));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += p[i] + l[i] + s[i] + a[i];
    if (r == 0x12345678u) out[0] = r;
}

template <int MODE>
void run_pipe(const char* name, int ops_per_iter, uint32_t* d_out, int nsm, int clk_khz) {
    int blocks = nsm * 8;
    pipe_kernel<MODE><<<blocks, 256>>>(d_out, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        pipe_kernel<MODE><<<blocks, 256>>>(d_out, 12345u + r);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double lane_ops = (double)blocks * 256 * ITERS * ops_per_iter;
    double per_s = lane_ops / (best * 1e-3);
    printf("PIPE %-28s ms=%8.3f lane_ops/s=%.4e  per_SM_per_clk@max=%.2f\n", name, best, per_s,
           per_s / nsm / (clk_khz * 1e3));
}

// ---- HBM streaming read ----
template <int UNROLL, bool NC>
__global__ void __launch_bounds__(256) stream_kernel(const uint4* __restrict__ in, size_t n_vec, unsigned long long* out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (; i + (UNROLL - 1) * stride < n_vec; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint4* p = in + i + u * stride;
            if (NC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
            else v[u] = *p;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc += __popc(v[u].x) + __popc(v[u].y) + __popc(v[u].z) + __popc(v[u].w);
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

template <int UNROLL, bool NC>
void run_stream(const char* name, const uint4* d, size_t bytes, unsigned long long* d_out, int nsm, int cta_per_sm) {
    size_t n_vec = bytes / 16;
    int blocks = nsm * cta_per_sm;
    stream_kernel<UNROLL, NC><<<blocks, 256>>>(d, n_vec, d_out);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < 10; r++) {
        CK(cudaEventRecord(e0));
        stream_kernel<UNROLL, NC><<<blocks, 256>>>(d, n_vec, d_out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; sum += ms;
    }
    printf("STREAM %-24s cta/sm=%d unroll=%d best_ms=%.4f avg_ms=%.4f best_GB/s=%.1f avg_GB/s=%.1f\n", name, cta_per_sm, UNROLL,
           best, sum / 10, bytes / (best * 1e-3) / 1e9, bytes / (sum / 10 * 1e-3) / 1e9);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount; int clk = prop.clockRate;
    printf("device=%s sms=%d clockRate_kHz=%d l2=%d MB\n", prop.name, nsm, clk, prop.l2CacheSize >> 20);
    uint32_t* d_out; CK(cudaMalloc(&d_out, 1024));
    run_pipe<0>("popc x8", 8, d_out, nsm, clk);
    run_pipe<1>("lop3 x8", 8, d_out, nsm, clk);
    run_pipe<2>("iadd x8", 8, d_out, nsm, clk);
    run_pipe<7>("imad x8", 8, d_out, nsm, clk);
    run_pipe<3>("popc x8 + lop3 x8", 16, d_out, nsm, clk);
    run_pipe<4>("popc x8 + lop3 x16", 24, d_out, nsm, clk);
    run_pipe<6>("popc x8 + lop3 x24", 32, d_out, nsm, clk);
    run_pipe<5>("popc x8 + lop3 x8 + iadd x8", 24, d_out, nsm, clk);
    run_pipe<8>("popc x8 + lop3 x8 + imad x8", 24, d_out, nsm, clk);
    run_pipe<9>("popc x4 + lop3 x16", 20, d_out, nsm, clk);
    run_pipe<10>("popc x8 + lop3 x4", 12, d_out, nsm, clk);

    size_t bytes = (size_t)4 << 30;
    uint4* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 0x5a, bytes));
    unsigned long long* d_o2; CK(cudaMalloc(&d_o2, 64));
    run_stream<1, false>("ldg128", d, bytes, d_o2, nsm, 8);
    run_stream<2, false>("ldg128", d, bytes, d_o2, nsm, 8);
    run_stream<4, false>("ldg128", d, bytes, d_o2, nsm, 8);
    run_stream<8, false>("ldg128", d, bytes, d_o2, nsm, 8);
    run_stream<4, true>("ldg128.nc.noalloc", d, bytes, d_o2, nsm, 8);
    run_stream<8, true>("ldg128.nc.noalloc", d, bytes, d_o2, nsm, 8);
    run_stream<4, true>("ldg128.nc.noalloc", d, bytes, d_o2, nsm, 4);
    run_stream<8, true>("ldg128.nc.noalloc", d, bytes, d_o2, nsm, 4);
    run_stream<8, true>("ldg128.nc.noalloc", d, bytes, d_o2, nsm, 2);
    run_stream<4, true>("ldg128.nc.noalloc", d, bytes, d_o2, nsm, 16);
    // cudaMemcpy D2D reference (read+write)
    {
        uint4* d2; CK(cudaMalloc(&d2, bytes));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMemcpy(d2, d, bytes, cudaMemcpyDeviceToDevice));
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("MEMCPY_D2D bytes_rw=%.1f GB best_ms=%.3f GB/s(read+write)=%.1f\n", 2.0 * bytes / 1e9, best, 2.0 * bytes / (best * 1e-3) / 1e9);
    }
    return 0;
}
