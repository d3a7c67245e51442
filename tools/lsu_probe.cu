// Micro-benchmark behind DESIGN.md's long-horizon analysis: the T = P22 L21 loop of invert_factor (n = 180) run by 3 lone
// warps per SM (one per scheduler), as in the H = 30 class.  Variants: V0 plain unrolled loop, V2 branch-free double buffering
// (what the kernel uses), V3 row-vector loads with multiply-adds paired over k (needs LD % 8 == 4), V4 / V5 a bank-conflicted
// and a conflict-free scalar sweep.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 [-DLD=188] -o lsu_probe lsu_probe.cu
// Result on B200 (profiles/r02_lsu_probe.txt): a lone warp pays 4 cycles of the shared-memory pipe per load instruction
// (32 unique bytes per clock and scheduler; a broadcast 128-bit load also 4, a 128-bit load of 32 distinct vectors 16), so the
// 2 x 4 register tile costs 12 load cycles + 8 multiply-add cycles per k and vector loads of the own rows buy nothing.
#include <cstdio>
#include <cuda_runtime.h>
#define N 180
#ifndef LD
#define LD 181
#endif
#define NT 96
__device__ __forceinline__ void load4(const float* p, float& a, float& b, float& c, float& d) { float4 v = *(const float4*)p; a = v.x; b = v.y; c = v.z; d = v.w; }
__device__ __forceinline__ void fma4_sub(float (&a)[4], float l, const float (&p)[4]) {
    const float2 nl = make_float2(-l, -l);
    const float2 r01 = __ffma2_rn(nl, make_float2(p[0], p[1]), make_float2(a[0], a[1]));
    const float2 r23 = __ffma2_rn(nl, make_float2(p[2], p[3]), make_float2(a[2], a[3]));
    a[0] = r01.x; a[1] = r01.y; a[2] = r23.x; a[3] = r23.y;
}
template <int V>
__global__ void __launch_bounds__(NT) probe(float* out, long long* cyc) {
    extern __shared__ float sm[];
    float* P = sm; float* stage = sm + ((LD * LD + 3) & ~3);
    const int tid = threadIdx.x;
    for (int i = tid; i < LD * LD; i += NT) P[i] = 1e-3f * (float)((i * 7) % 13);
    for (int i = tid; i < 4 * 192; i += NT) stage[i] = 1e-3f * (float)(i % 5);
    __syncthreads();
    float tot = 0;
    long long t0 = clock64();
    for (int j0 = N - 4; j0 >= 0; j0 -= 4) {
        float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        const int m0 = (j0 + 4) / NT;
        const float* pc = P + (size_t)(j0 + 4) * LD + tid;
        const float* sg = stage + 4 * (j0 + 4);
        if (V == 0) {
#pragma unroll 2
            for (int k = j0 + 4; k < N; k += 4) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float l[4];
                    load4(sg + 4 * t, l[0], l[1], l[2], l[3]);
#pragma unroll
                    for (int m = 0; m < 2; ++m) { if (m < m0) continue; fma4_sub(acc[m], pc[t * LD + NT * m], l); }
                }
                pc += 4 * LD; sg += 16;
            }
        } else if (V == 1) {
            // explicit software pipeline, one k ahead, volatile asm keeps the order LDS, FFMA2, FFMA2 ...
            unsigned pa = (unsigned)__cvta_generic_to_shared(pc), sa = (unsigned)__cvta_generic_to_shared(sg);
            float l0, l1, l2, l3, a0, a1;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l0), "=f"(l1), "=f"(l2), "=f"(l3) : "r"(sa));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(pa));
            asm volatile("ld.shared.f32 %0, [%1+384];" : "=f"(a1) : "r"(pa));
            unsigned long long c00, c01, c10, c11;   // packed accumulators
            asm volatile("mov.b64 %0, 0; mov.b64 %1, 0; mov.b64 %2, 0; mov.b64 %3, 0;" : "=l"(c00), "=l"(c01), "=l"(c10), "=l"(c11));
            for (int k = j0 + 4; k < N; ++k) {
                float n0, n1, n2, n3, b0, b1;
                sa += 16; pa += 4 * LD;
                unsigned long long la, lb, m0_, m1_;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(la) : "f"(l0), "f"(l1));
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(lb) : "f"(l2), "f"(l3));
                float na0 = -a0, na1 = -a1;
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(m0_) : "f"(na0));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(m1_) : "f"(na1));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(n0), "=f"(n1), "=f"(n2), "=f"(n3) : "r"(sa));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c00) : "l"(m0_), "l"(la));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c01) : "l"(m0_), "l"(lb));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b0) : "r"(pa));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c10) : "l"(m1_), "l"(la));
                asm volatile("ld.shared.f32 %0, [%1+384];" : "=f"(b1) : "r"(pa));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c11) : "l"(m1_), "l"(lb));
                l0 = n0; l1 = n1; l2 = n2; l3 = n3; a0 = b0; a1 = b1;
            }
            asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[0][0]), "=f"(acc[0][1]) : "l"(c00));
            asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[0][2]), "=f"(acc[0][3]) : "l"(c01));
            asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[1][0]), "=f"(acc[1][1]) : "l"(c10));
            asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(acc[1][2]), "=f"(acc[1][3]) : "l"(c11));
        } else if (V == 2) {
            // branch-free double-buffered groups of 4 k: fetch(g+1) and apply(g) are independent inside one basic block
            float la[4][4], pa[4][2], lb[4][4], pb[4][2];
            auto fetch = [&](float (&l)[4][4], float (&p)[4][2], const float* pcc, const float* sgg) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    load4(sgg + 4 * t, l[t][0], l[t][1], l[t][2], l[t][3]);
                    p[t][0] = pcc[t * LD]; p[t][1] = pcc[t * LD + NT];
                }
            };
            auto apply = [&](const float (&l)[4][4], const float (&p)[4][2]) {
#pragma unroll
                for (int t = 0; t < 4; ++t) { fma4_sub(acc[0], p[t][0], l[t]); fma4_sub(acc[1], p[t][1], l[t]); }
            };
            const int ng = (N - j0 - 4) >> 2;
            fetch(la, pa, pc, sg);
            int g = 0;
#pragma unroll 1
            for (; g + 1 < ng; g += 2) {
                fetch(lb, pb, pc + 4 * LD, sg + 16);
                apply(la, pa);
                pc += 8 * LD; sg += 32;
                fetch(la, pa, pc, sg);                 // may read one group past the end: harmless
                apply(lb, pb);
            }
            if (g < ng) apply(la, pa);
        } else if (V == 3) {
            // row-vector loads (needs LD % 8 == 4): FFMA2 pairs over k; stage holds L21 transposed: stageT[c][k]
            float2 a2[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c) a2[m][c] = make_float2(0.f, 0.f);
            const float* pr0 = P + (size_t)tid * LD + (j0 + 4);
            const float* pr1 = pr0 + NT * LD;
            const float* st = stage + (j0 + 4);
#pragma unroll 2
            for (int k = j0 + 4; k < N; k += 4) {
                float4 p0 = *(const float4*)pr0, p1 = *(const float4*)pr1;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 l = *(const float4*)(st + c * 192);
                    a2[0][c] = __ffma2_rn(make_float2(p0.x, p0.y), make_float2(l.x, l.y), a2[0][c]);
                    a2[0][c] = __ffma2_rn(make_float2(p0.z, p0.w), make_float2(l.z, l.w), a2[0][c]);
                    a2[1][c] = __ffma2_rn(make_float2(p1.x, p1.y), make_float2(l.x, l.y), a2[1][c]);
                    a2[1][c] = __ffma2_rn(make_float2(p1.z, p1.w), make_float2(l.z, l.w), a2[1][c]);
                }
                pr0 += 4; pr1 += 4; st += 4;
            }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[m][c] = a2[m][c].x + a2[m][c].y;
        } else if (V == 4) {
            // cost of the transposed scalar read (bank-conflicted when LD % 8 == 4): one load per k, no arithmetic to speak of
            const float* pr0 = P + (size_t)tid * LD + (j0 + 4);
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 2
            for (int k = j0 + 4; k < N; k += 4) { s0 += pr0[0]; s1 += pr0[1]; s2 += pr0[2]; s3 += pr0[3]; pr0 += 4; }
            acc[0][0] = s0 + s1 + s2 + s3;
        } else if (V == 5) {
            const float* pc0 = P + (size_t)(j0 + 4) * LD + tid;
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 2
            for (int k = j0 + 4; k < N; k += 4) { s0 += pc0[0]; s1 += pc0[LD]; s2 += pc0[2 * LD]; s3 += pc0[3 * LD]; pc0 += 4 * LD; }
            acc[0][0] = s0 + s1 + s2 + s3;
        }
        tot += acc[0][0] + acc[0][1] + acc[0][2] + acc[0][3] + acc[1][0] + acc[1][1] + acc[1][2] + acc[1][3];
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * NT + tid] = tot;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * NT * 4 * 8); cudaMalloc(&cyc, 148 * 8 * 8);
    size_t smem = (196 * LD + 4 * 192 + 16 + 8 * LD) * 4;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[148];
#define RUN(V) { cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); for (int rep = 0; rep < 2; ++rep) { probe<V><<<148, NT, smem>>>(out, cyc); cudaDeviceSynchronize(); } cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); printf("LD %d V%d cycles %lld  per k %.2f (%s)\n", LD, V, h[0], h[0] / 3960.0, cudaGetErrorString(cudaGetLastError())); }
    RUN(0) RUN(2) RUN(3) RUN(4) RUN(5)
    float ho[4]; cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost); printf("%g\n", ho[0]);
    return 0;
}
