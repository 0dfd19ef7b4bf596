// Memory-pattern micro-benchmarks on one GPU (development tool, not part of the product):
// what the hardware gives for the access patterns of the remap kernel, measured with CUDA events.
//   fill (write-only), sum (read-only), copy, row gather (640-byte rows at random / sorted
//   positions out of a 2.36 GB field, read-only), row gather + row store (the remap shape
//   without arithmetic or reuse).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void ld256(const double *p, double *v) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
template <int CS>
__device__ __forceinline__ void st256(double *p, const double *v) {
    if (CS) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

template <int CS>
__global__ void fill_kernel(double *y, long long n4) {
    double v[4] = {1.0, 2.0, 3.0, 4.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        st256<CS>(y + i * 4, v);
}
__global__ void sum_kernel(const double *x, long long n4, double *out) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        double v[4];
        ld256(x + i * 4, v);
        acc += v[0] + v[1] + v[2] + v[3];
    }
    if (acc == 12345.678) *out = acc;
}
template <int CS>
__global__ void copy_kernel(const double *x, double *y, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        double v[4];
        ld256(x + i * 4, v);
        st256<CS>(y + i * 4, v);
    }
}
// one 640-byte row per 20 lanes (blockDim = (20, 8) like the PBIN kernel); UNROLL rows in flight per thread
template <int UNROLL, int STORE>
__global__ void gather_rows(const double *x, const int *rows, int n_rows, double *y, double *out) {
    const int lx = threadIdx.x;
    double acc = 0.0;
    const int per = blockDim.y * UNROLL;
    for (long long base = (long long)blockIdx.x * per; base < n_rows; base += (long long)gridDim.x * per) {
        double v[UNROLL][4];
        int r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long i = base + u * blockDim.y + threadIdx.y;
            r[u] = i < n_rows ? rows[i] : -1;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (r[u] >= 0) ld256(x + (long long)r[u] * 80 + lx * 4, v[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (r[u] < 0) continue;
            if (STORE) {
                const long long i = base + u * blockDim.y + threadIdx.y;
                st256<1>(y + i * 80 + lx * 4, v[u]);
            } else {
                acc += v[u][0] + v[u][1] + v[u][2] + v[u][3];
            }
        }
    }
    if (!STORE && acc == 12345.678) *out = acc;
}
// same rows, but a warp takes 8 rows x 128 bytes per pass (the WROW shape), 5 passes
template <int STORE>
__global__ void gather_rows_w(const double *x, const int *rows, int n_rows, double *y, double *out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (long long base = warp * 8; base < n_rows; base += n_warps * 8) {
        const long long i = base + g;
        const int r = i < n_rows ? rows[i] : -1;
        if (r < 0) continue;
        double v[5][4];
#pragma unroll
        for (int p = 0; p < 5; ++p) ld256(x + (long long)r * 80 + (p * 4 + c) * 4, v[p]);
#pragma unroll
        for (int p = 0; p < 5; ++p) {
            if (STORE) st256<1>(y + i * 80 + (p * 4 + c) * 4, v[p]);
            else acc += v[p][0] + v[p][1] + v[p][2] + v[p][3];
        }
    }
    if (!STORE && acc == 12345.678) *out = acc;
}

template <typename F>
float time_it(F f, int reps = 10) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}
void report(const char *what, double bytes, float ms) {
    printf("%-64s %9.1f us  %8.1f GB/s\n", what, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
}

int main() {
    const long long n_cells = 3693225, K = 80;
    const long long n_field = n_cells * K;            // doubles per slice (2.36 GB)
    const int slices = 4;
    double *x, *y, *out;
    CK(cudaMalloc(&x, sizeof(double) * n_field * slices));
    CK(cudaMalloc(&y, sizeof(double) * n_field * 2));
    CK(cudaMalloc(&out, 8));
    CK(cudaMemset(x, 0, sizeof(double) * n_field * slices));
    int sms = 148;
    const int grid = sms * 8;
    // streams
    {
        const long long n4 = n_field * 2 / 4;     // 4.7 GB
        report("fill  st.global      4.7 GB", n4 * 32.0, time_it([&] { fill_kernel<0><<<grid, 256>>>(y, n4); }));
        report("fill  st.global.cs   4.7 GB", n4 * 32.0, time_it([&] { fill_kernel<1><<<grid, 256>>>(y, n4); }));
        report("cudaMemsetAsync      4.7 GB", n4 * 32.0, time_it([&] { CK(cudaMemsetAsync(y, 0, n4 * 32)); }));
        report("sum   ld.global.nc   4.7 GB", n4 * 32.0, time_it([&] { sum_kernel<<<grid, 256>>>(x, n4, out); }));
        report("copy  ld.nc/st       4.7+4.7 GB", n4 * 64.0, time_it([&] { copy_kernel<0><<<grid, 256>>>(x, y, n4); }));
        report("copy  ld.nc/st.cs    4.7+4.7 GB", n4 * 64.0, time_it([&] { copy_kernel<1><<<grid, 256>>>(x, y, n4); }));
        const long long s4 = 193000000 / 32;      // one output slice
        report("fill  st.global.cs   193 MB (one Y slice)", s4 * 32.0, time_it([&] { fill_kernel<1><<<grid, 256>>>(y, s4); }));
        const long long b4 = 8 * s4;
        report("fill  st.global.cs   1.54 GB (8 Y slices)", b4 * 32.0, time_it([&] { fill_kernel<1><<<grid, 256>>>(y, b4); }));
    }
    // row gathers: 337090 distinct rows out of 3.69M (the C3 touched set is a compact band; here a
    // random sample, sorted or shuffled), 8 slices per launch -> rows index into 4 slices
    std::mt19937_64 rng(7);
    for (int pattern = 0; pattern < 3; ++pattern) {
        const int n_rows = 337090 * 8;
        std::vector<int> rows(n_rows);
        for (int s = 0; s < 8; ++s) {
            std::vector<int> pick(337090);
            if (pattern == 2) {          // compact band: consecutive rows starting at a random offset
                for (int i = 0; i < 337090; ++i) pick[i] = 1000000 + i;
            } else {
                for (int i = 0; i < 337090; ++i) pick[i] = (int)(rng() % n_cells);
                if (pattern == 0) std::sort(pick.begin(), pick.end());
            }
            for (int i = 0; i < 337090; ++i) rows[s * 337090 + i] = pick[i] + (s % slices) * (int)n_cells;
        }
        int *d_rows;
        CK(cudaMalloc(&d_rows, sizeof(int) * n_rows));
        CK(cudaMemcpy(d_rows, rows.data(), sizeof(int) * n_rows, cudaMemcpyHostToDevice));
        const char *pn = pattern == 0 ? "sorted random rows" : pattern == 1 ? "shuffled random rows" : "consecutive rows";
        char buf[128];
        const double rb = (double)n_rows * 640.0;
        dim3 blk(20, 8);
        snprintf(buf, sizeof buf, "gather 640B rows, %s, read-only, 1 in flight", pn);
        report(buf, rb, time_it([&] { gather_rows<1, 0><<<sms * 6, blk>>>(x, d_rows, n_rows, y, out); }));
        snprintf(buf, sizeof buf, "gather 640B rows, %s, read-only, 4 in flight", pn);
        report(buf, rb, time_it([&] { gather_rows<4, 0><<<sms * 6, blk>>>(x, d_rows, n_rows, y, out); }));
        snprintf(buf, sizeof buf, "gather 640B rows, %s, read-only, 8 in flight", pn);
        report(buf, rb, time_it([&] { gather_rows<8, 0><<<sms * 4, blk>>>(x, d_rows, n_rows, y, out); }));
        snprintf(buf, sizeof buf, "gather+store rows, %s, 4 in flight (r+w bytes)", pn);
        report(buf, 2 * rb, time_it([&] { gather_rows<4, 1><<<sms * 6, blk>>>(x, d_rows, n_rows, y, out); }));
        snprintf(buf, sizeof buf, "gather rows warp=8x128B x5, %s, read-only", pn);
        report(buf, rb, time_it([&] { gather_rows_w<0><<<sms * 4, 256>>>(x, d_rows, n_rows, y, out); }));
        snprintf(buf, sizeof buf, "gather+store warp=8x128B x5, %s (r+w bytes)", pn);
        report(buf, 2 * rb, time_it([&] { gather_rows_w<1><<<sms * 4, 256>>>(x, d_rows, n_rows, y, out); }));
        CK(cudaFree(d_rows));
    }
    return 0;
}
