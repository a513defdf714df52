// Pipe-throughput microbenchmark for the integer ops the 2048 kernels are made of (sm_100a).
// Each thread runs NCHAIN independent dependency chains of one op; 148 x k CTAs of 1024 threads.
// Reports warp-instructions per cycle per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCHAIN 8

enum Op { LOP3, SHF_R, SHL_IMAD, PRMT, IADD3, IMAD, IMAD_HI, IMAD_WIDE, ISETP_SEL, POPC, LEA, VIMNMX, MIX_LOP_IMAD, MIX_LOP_IMADHI, MIX_SHF_IMADSHL, LDS_NOCONF, LDS_RANDOM, IADD_IMAD, MIX_LOP_WIDE, MIX_IMM, MIX3, MIX_2TO1, LOP3_IMM, IMAD_IMM, MIX_IMM_2TO1, MIX_IMM_3TO1, MIX_IMM_5TO3, MIX_IMM_WIDE_3TO1, MIX_IMM_LDS, NOPS };
const char* names[] = {"LOP3","SHF.R","IMAD.SHL(mul 16)","PRMT","IADD3","IMAD","IMAD.HI.U32","IMAD.WIDE.U32","ISETP+SEL","POPC","LEA","VIMNMX","mix LOP3+IMAD 1:1","mix LOP3+IMAD.HI 1:1","mix SHF+IMAD.SHL 1:1","LDS no conflict","LDS random u16 idx","mad.lo x*1+y","mix LOP3+IMAD.WIDE 1:1","mix LOP3(imm)+IMAD(imm) 1:1","mix LOP3+IMAD+PRMT+IMAD","mix LOP3:IMAD 2:1","LOP3 (imm operands)","IMAD (imm operands)","mix LOP3:IMAD imm 2:1","mix LOP3:IMAD imm 3:1","mix LOP3:IMAD imm 5:3","mix LOP3(imm):IMAD.WIDE 3:1","mix LOP3:IMAD imm 1:1 + 1 LDS per 16"};

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, long long* cycles)
{
    __shared__ uint32_t sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t x[NCHAIN];
#pragma unroll
    for (int c = 0; c < NCHAIN; c++) x[c] = seed + threadIdx.x * 977u + c * 131u + blockIdx.x;
    uint32_t y = seed | 1u, z = seed * 3u + 7u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < NCHAIN; c++) {
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z));
            if (OP == SHF_R) asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(x[c]) : "r"(y));
            if (OP == SHL_IMAD) asm volatile("mul.lo.u32 %0, %0, 16;" : "+r"(x[c]));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2301;" : "+r"(x[c]) : "r"(y));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            if (OP == IMAD_HI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            if (OP == IMAD_WIDE) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(y)); x[c] = (uint32_t)p ^ (uint32_t)(p >> 32); }
            if (OP == ISETP_SEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x[c]) : "r"(y), "r"(z));
            if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            if (OP == LEA) asm volatile("{.reg .u32 t; shl.b32 t, %0, 2; add.u32 %0, t, %1;}" : "+r"(x[c]) : "r"(y));
            if (OP == VIMNMX) asm volatile("max.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            if (OP == MIX_LOP_IMAD) { if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); }
            if (OP == MIX_LOP_IMADHI) { if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); }
            if (OP == MIX_SHF_IMADSHL) { if (c & 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(x[c]) : "r"(y)); else asm volatile("mul.lo.u32 %0, %0, 16;" : "+r"(x[c])); }
            if (OP == LDS_NOCONF) { uint32_t a = sbase + (threadIdx.x & 31) * 4 + (x[c] & 0x7f00); asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[c]) : "r"(a)); }
            if (OP == LDS_RANDOM) { uint32_t a = sbase + (x[c] & 0x7ffe); uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); x[c] = x[c] * 5u + v; }
            if (OP == IADD_IMAD) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[c]) : "r"(y));
            if (OP == MIX_IMM) { if (c & 1) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c])); else asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 0x7F4A7C15;" : "+r"(x[c])); }
            if (OP == MIX3) { if ((c & 3) == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); else if ((c & 3) == 2) asm volatile("prmt.b32 %0, %0, %1, 0x2301;" : "+r"(x[c]) : "r"(y)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); }
            if (OP == MIX_2TO1) { if ((c % 3) != 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z)); }
            if (OP == LOP3_IMM) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c]));
            if (OP == IMAD_IMM) asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 0x7F4A7C15;" : "+r"(x[c]));
            if (OP == MIX_IMM_2TO1) { if ((c % 3) != 2) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c])); else asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 0x7F4A7C15;" : "+r"(x[c])); }
            if (OP == MIX_IMM_3TO1) { if ((c & 3) != 3) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c])); else asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 0x7F4A7C15;" : "+r"(x[c])); }
            if (OP == MIX_IMM_5TO3) { if ((c & 7) < 5) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c])); else asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 0x7F4A7C15;" : "+r"(x[c])); }
            if (OP == MIX_IMM_WIDE_3TO1) { if ((c & 3) != 3) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c])); else { uint64_t p; asm volatile("mul.wide.u32 %0, %1, 0x9E3779B1;" : "=l"(p) : "r"(x[c])); x[c] = (uint32_t)(p >> 32) ^ (uint32_t)p; } }
            if (OP == MIX_IMM_LDS) { if (c & 1) asm volatile("lop3.b32 %0, %0, 0x0f0f0f0f, 0x12345678, 0x96;" : "+r"(x[c])); else asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 0x7F4A7C15;" : "+r"(x[c])); if (c == 0 && (it & 1) == 0) { uint32_t a = sbase + (threadIdx.x & 31) * 4 + (x[1] & 0x7f00); uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); x[7] ^= v; } }
            if (OP == MIX_LOP_WIDE) { if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z)); else { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(y)); x[c] = (uint32_t)(p >> 32) + (uint32_t)p; } }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < NCHAIN; c++) acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(uint32_t* out, long long* cyc, int sms, double extra_per_iter)
{
    k<OP><<<sms, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<sms, 1024>>>(out, 12345u, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double winstr = 32.0 * ITERS * NCHAIN;          // warp-instr of the op per SM (32 warps)
    printf("%-26s %8.3f target-op warp-instr/cycle/SM   (%.0f cycles, %.3f ms)\n", names[OP], winstr / avg, avg, ms);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&cyc, 1024 * 8);
    printf("SMs %d; ops counted as written in PTX (check SASS for what ptxas emitted)\n", sms);
    run<LOP3>(out, cyc, sms, 0); run<SHF_R>(out, cyc, sms, 0); run<SHL_IMAD>(out, cyc, sms, 0); run<PRMT>(out, cyc, sms, 0);
    run<IADD3>(out, cyc, sms, 0); run<IMAD>(out, cyc, sms, 0); run<IMAD_HI>(out, cyc, sms, 0); run<IMAD_WIDE>(out, cyc, sms, 0);
    run<ISETP_SEL>(out, cyc, sms, 0); run<POPC>(out, cyc, sms, 0); run<LEA>(out, cyc, sms, 0); run<VIMNMX>(out, cyc, sms, 0);
    run<MIX_LOP_IMAD>(out, cyc, sms, 0); run<MIX_LOP_IMADHI>(out, cyc, sms, 0); run<MIX_SHF_IMADSHL>(out, cyc, sms, 0);
    run<LDS_NOCONF>(out, cyc, sms, 0); run<LDS_RANDOM>(out, cyc, sms, 0); run<IADD_IMAD>(out, cyc, sms, 0); run<MIX_LOP_WIDE>(out, cyc, sms, 0); run<MIX_IMM>(out, cyc, sms, 0); run<MIX3>(out, cyc, sms, 0); run<MIX_2TO1>(out, cyc, sms, 0);
    run<LOP3_IMM>(out, cyc, sms, 0); run<IMAD_IMM>(out, cyc, sms, 0); run<MIX_IMM_2TO1>(out, cyc, sms, 0); run<MIX_IMM_3TO1>(out, cyc, sms, 0); run<MIX_IMM_5TO3>(out, cyc, sms, 0); run<MIX_IMM_WIDE_3TO1>(out, cyc, sms, 0); run<MIX_IMM_LDS>(out, cyc, sms, 0);
    return 0;
}
