// Issue-rate microbenchmark for the instruction classes the detection kernels are made of (sm_100a).
// Each test: 8 independent chains per thread, ITER iterations, 16 warps per SM (4 per scheduler), one CTA per SM.
// Prints warp-instructions per cycle per SM sub-partition (1.0 = the issue limit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int T>
__device__ __forceinline__ void body(uint32_t (&r)[8], float (&f)[8], unsigned long long (&d)[8], uint32_t k, float fk, unsigned long long dk, const uint32_t* sm)
{
	const uint32_t smb = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
	for (int i = 0; i < 8; i++) {
		if (T == 0) asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k));
		if (T == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fk));
		if (T == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk));
		if (T == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dk));
		if (T == 4) { float t; asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(t) : "r"(r[i] >> 8)); r[i] = __float_as_uint(t); } // I2F.U8 Rx.B1, chained through the bits
		if (T == 5) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(r[i]) : "r"(k));
		if (T == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(r[i]) : "r"(k));
		if (T == 7) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(k));
		if (T == 8) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(r[i]));
		if (T == 9) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fk));
		if (T == 10) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r[i]) : "r"(smb + r[i]));
		if (T == 11) { uint32_t a, b; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(smb + 2 * r[i])); r[i] = a; }
		if (T == 12) { uint32_t a, b, c, e; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(e) : "r"(smb + 4 * r[i])); r[i] = a; }
		if (T == 13) { // PRMT + FFMA2 alternating (the blend's core pair)
			asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k));
			asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk));
		}
		if (T == 14) { // 2 PRMT + 1 FFMA2 (the blend's real ratio)
			asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k));
			asm volatile("prmt.b32 %0, %0, %1, 0x7441;" : "+r"(r[i]) : "r"(k));
			asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk));
		}
		if (T == 24) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk)); asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fk)); }
		if (T == 25) { asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fk)); }
		if (T == 26) { asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(k)); asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[(i + 4) & 7]) : "r"(k)); }
		if (T == 27) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fk)); }
		if (T == 28) asm volatile("shfl.sync.down.b32 %0, %0, 3, 31, 0xffffffff;" : "+r"(r[i]));
		if (T == 29) { asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(r[i])); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(fk)); asm volatile("min.f32 %0, %0, %1;" : "+f"(f[(i + 4) & 7]) : "f"(fk)); }
		if (T == 30) asm volatile("lea.hi.u32 %0, %0, %1, %1;" :: "r"(r[i]), "r"(k)); /* placeholder, not run */
		if (T == 31) { asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k)); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dk)); }
		if (T == 32) { float t; asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(t) : "r"(r[i] >> 8)); r[i] = __float_as_uint(t); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk)); }
		if (T == 33) { asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(dk), "l"(d[(i + 1) & 7])); }
		if (T == 34) asm volatile("st.shared.b32 [%0], %1;" :: "r"(smb + ((r[i] + i * 4) & 8188)), "r"(r[i]));
		if (T == 35) { float t; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(r[i]) : "f"(f[i])); asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(r[i])); f[i] = t; }
		if (T == 15) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fk));
		if (T == 16) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dk));
		if (T == 17) { // denormal operand: FMUL2 on (b * 2^-149) values
			asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dk)); /* d holds denormals (set up by the caller), dk = (1, 1) */
		}
		if (T == 18) { // I2F.U8 + PRMT + FFMA2 three-pipe mix
			{ float t; asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(t) : "r"(r[(i + 4) & 7] >> 8)); f[i] = t; }
			asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(r[i]) : "r"(k));
			asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dk));
		}
		if (T == 19) { // HADD2.F32 conversion of an fp16 half
			float t; asm volatile("{ .reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi; }" : "=f"(t) : "r"(r[i])); r[i] = __float_as_uint(t);
		}
		if (T == 20) asm volatile("dp4a.u32.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(k));
		if (T == 21) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(k));
		if (T == 22) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(f[i]) : "f"(fk)); // immediate multiplier
		if (T == 23) { // FMNMX + FFMA alternating
			asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fk));
			asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[(i + 4) & 7]) : "f"(fk));
		}
	}
}

template <int T>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, uint32_t k0, float fk, unsigned long long dk, long long* cycles)
{
	__shared__ uint32_t sm[2048];
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = T == 10 ? (uint32_t)(((i + 32) & 2047) * 4) : T == 11 ? (uint32_t)((((i >> 1) + 32) & 511) * 4) : (uint32_t)((((i >> 2) + 32) & 255) * 4);
	__syncthreads();
	uint32_t r[8];
	float f[8];
	unsigned long long d[8];
#pragma unroll
	for (int i = 0; i < 8; i++) { r[i] = (T >= 10 && T <= 12) ? (uint32_t)(((threadIdx.x & 31) + 32 * i) * 4) : threadIdx.x * 77 + i + k0; f[i] = (float)r[i] * 1e-3f; d[i] = T == 17 ? (((unsigned long long)(i + 3) << 32) | (threadIdx.x & 255)) : (((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f)); }
	__syncthreads();
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITER / 8; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) body<T>(r, f, d, k0, fk, dk, sm);
	}
	const long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) acc ^= r[i] ^ __float_as_uint(f[i]) ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32);
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int T>
void run(const char* name, int per_iter, uint32_t* out, long long* cyc)
{
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	float one = 1.0000001f;
	unsigned long long dk = T == 17 ? (((unsigned long long)0x3F800000u << 32) | 0x3F800000u) : (((unsigned long long)0x3F800001u << 32) | 0x3F800001u);
	k<T><<<sms, 512>>>(out, 0x4B000000u, one, dk, cyc);
	k<T><<<sms, 512>>>(out, 0x4B000000u, one, dk, cyc);
	cudaDeviceSynchronize();
	long long h[1024];
	cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < sms; i++) avg += (double)h[i];
	avg /= sms;
	// 16 warps per SM = 4 per sub-partition; each executes ITER * per_iter instructions of the class
	const double ipc = 4.0 * ITER * per_iter / avg;
	printf("%-34s %8.0f cycles   %.3f warp-instr/clk/SMSP   (rt %.2f)\n", name, avg, ipc, 1.0 / ipc);
}

int main()
{
	uint32_t* out;
	long long* cyc;
	cudaMalloc(&out, 148 * 2 * 512 * 4);
	cudaMalloc(&cyc, 1024 * 8);
	run<0>("PRMT", 8, out, cyc);
	run<1>("FFMA (3 reg)", 8, out, cyc);
	run<22>("FFMA (imm)", 8, out, cyc);
	run<15>("FADD", 8, out, cyc);
	run<2>("FFMA2", 8, out, cyc);
	run<3>("FMUL2", 8, out, cyc);
	run<16>("FADD2", 8, out, cyc);
	run<17>("FMUL2 denormal operand", 8, out, cyc);
	run<4>("I2F.U8 (byte select)", 8, out, cyc);
	run<19>("HADD2.F32 (f16 -> f32)", 8, out, cyc);
	run<5>("LOP3", 8, out, cyc);
	run<6>("SHF", 8, out, cyc);
	run<21>("IADD", 8, out, cyc);
	run<7>("IMAD", 8, out, cyc);
	run<20>("DP4A", 8, out, cyc);
	run<9>("FMNMX", 8, out, cyc);
	run<8>("SHFL", 8, out, cyc);
	run<10>("LDS.32", 8, out, cyc);
	run<11>("LDS.64", 8, out, cyc);
	run<12>("LDS.128", 8, out, cyc);
	run<13>("PRMT + FFMA2 (1:1)", 16, out, cyc);
	run<14>("2 PRMT + FFMA2", 24, out, cyc);
	run<18>("I2F.U8 + PRMT + FFMA2", 24, out, cyc);
	run<23>("FMNMX + FFMA (1:1)", 16, out, cyc);
	run<24>("FFMA2 + FMNMX (1:1)", 16, out, cyc);
	run<25>("PRMT + FFMA (1:1)", 16, out, cyc);
	run<26>("IMAD + PRMT (1:1)", 16, out, cyc);
	run<27>("FFMA2 + FFMA (1:1)", 16, out, cyc);
	run<31>("PRMT + FMUL2 (1:1)", 16, out, cyc);
	run<33>("PRMT + FFMA2 3 distinct operands", 16, out, cyc);
	run<32>("I2F.U8 + FFMA2 (1:1)", 16, out, cyc);
	run<28>("SHFL.DOWN", 8, out, cyc);
	run<29>("SHFL + FFMA + FMNMX", 24, out, cyc);
	run<34>("STS.32", 8, out, cyc);
	run<35>("F2I + I2F", 16, out, cyc);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
