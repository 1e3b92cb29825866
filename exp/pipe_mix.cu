// Micro-benchmark: issue rates of single instruction forms and of two-pipe mixtures on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp/pipe_mix exp/pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CH = 8;

template <int W>
__global__ void __launch_bounds__(256) k(float fs, double ds, unsigned us, unsigned* sink) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    float f[CH], g[CH]; double d[CH]; unsigned u[CH], v[CH]; unsigned long long w[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { f[c] = 1.f + fs * (tid + c); g[c] = 0.5f + fs * c; d[c] = 1.0 + ds * (tid + c); u[c] = tid * 2654435761u + c; v[c] = u[c] ^ us; w[c] = u[c]; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (W == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(g[c]), "f"(fs));           // FFMA rrr
            if (W == 1) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(f[c]) : "f"(g[c]));            // FFMA rr,imm
            if (W == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(g[c]));                       // FMUL rr
            if (W == 3) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(f[c]));                           // FMUL r,imm
            if (W == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(us));      // LOP3 rrr(UR)
            if (W == 5) asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]));  // IMAD.WIDE
            if (W == 6) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(ds));                     // DFMA
            if (W == 7) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));                               // MUFU
            // mixtures: one of each per chain step
            if (W == 10) { asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(us)); }
            if (W == 11) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(ds));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(us)); }
            if (W == 12) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(ds));
                           asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(g[c]), "f"(fs)); }
            if (W == 13) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(g[c]), "f"(fs));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(us)); }
            if (W == 14) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(g[c]));
                           asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fs), "f"(fs)); }
            if (W == 15) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(ds));
                           asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[c]) : "r"((unsigned)w[c])); }
            if (W == 16) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(ds));
                           asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[c]) : "r"((unsigned)w[c]));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(us)); }
            if (W == 17) { asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(f[c]) : "f"(g[c]));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(us)); }
            if (W == 18) { asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(f[c]) : "f"(g[c]));
                           asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[c]) : "r"((unsigned)w[c])); }
            if (W == 19) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[c]) : "f"(f[c]));                      // F2F alone (plus the chain below)
                           asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(f[c]) : "f"(g[c])); }
            if (W == 20) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[c]) : "r"(u[c]));                   // I2FP + LOP3
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(__float_as_uint(f[c]))); }
            if (W == 21) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[c]) : "r"(u[c]));                   // I2FP + IMAD.WIDE
                           asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[c]) : "r"(__float_as_uint(f[c]))); u[c] = (unsigned)w[c]; }
        }
    }
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= __float_as_uint(f[c]) ^ __float_as_uint(g[c]) ^ (unsigned)d[c] ^ u[c] ^ (unsigned)w[c];
    if (acc == 0x12345u) sink[0] = acc;
}

template <int W> void run(const char* name, int per_step) {
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * 8;
    k<W><<<grid, 256>>>(1e-7f, 1e-12, 3u, sink);
    cudaEventRecord(a); k<W><<<grid, 256>>>(1e-7f, 1e-12, 3u, sink); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double warp_inst = (double)grid * 8 * ITERS * CH * per_step;
    // SMSP-cycles available = 148*4*f*t ; report instructions per SMSP cycle assuming 1.965 GHz
    double ipc = warp_inst / (148.0 * 4 * 1.965e9 * ms * 1e-3);
    printf("%-34s %8.3f ms  %6.3f warp-inst/clk/SMSP (at 1965 MHz)   %.2f cyc per chain step\n", name, ms, ipc, per_step / ipc);
    cudaFree(sink);
}

int main() {
    run<0>("FFMA r,r,r", 1); run<1>("FFMA r,r,imm", 1); run<2>("FMUL r,r", 1); run<3>("FMUL r,imm", 1);
    run<4>("LOP3 r,r,ur", 1); run<5>("IMAD.WIDE r,imm,r", 1); run<6>("DFMA", 1); run<7>("MUFU.LG2", 1);
    run<10>("IMAD.WIDE + LOP3", 2); run<11>("DFMA + LOP3", 2); run<12>("DFMA + FFMA rrr", 2); run<13>("FFMA rrr + LOP3", 2);
    run<14>("MUFU + FFMA", 2); run<15>("DFMA + IMAD.WIDE", 2); run<16>("DFMA + IMAD.WIDE + LOP3", 3);
    run<17>("FFMA imm + LOP3", 2); run<18>("FFMA imm + IMAD.WIDE", 2); run<19>("F2F.F64.F32 + FFMA imm", 2);
    run<20>("I2FP + LOP3", 2); run<21>("I2FP + IMAD.WIDE", 2);
    return 0;
}
