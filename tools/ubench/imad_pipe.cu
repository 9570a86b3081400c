// Micro-benchmark (development aid): issue cost of the limb-product instruction forms on sm_100a.
//   A  mad.wide.u32 acc64 += a*b                (IMAD.WIDE.U32, no carry)
//   B  mad.lo.cc / madc.hi.cc chains            (IMAD.WIDE.U32.X, carry in + out)
//   C  mul.wide.u32 + add.cc/addc.cc            (IMAD.WIDE.U32 fresh + 2 IADD3.X)
//   D  mad.lo.u32                               (IMAD 32-bit)
//   E  B interleaved with independent IADD3     (does the alu pipe issue alongside?)
// nvcc -arch=sm_100a -O3 -o imad_pipe imad_pipe.cu && ./imad_pipe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int NCH = 8;  // limbs per chain row

template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed)
{
	uint32_t a[NCH], acc[NCH + 2], x = threadIdx.x * 2654435761u + seed, y = seed ^ 0x9e3779b9u, z = 0;
#pragma unroll
	for (int j = 0; j < NCH; ++j) a[j] = x * (j + 3) + 1;
#pragma unroll
	for (int j = 0; j < NCH + 2; ++j) acc[j] = j;
#pragma unroll 1
	for (int it = 0; it < ITER; ++it)
	{
		if (V == 0)
		{
#pragma unroll
			for (int j = 0; j < NCH; j += 2)
			{
				uint64_t t = (uint64_t(acc[j + 1]) << 32) | acc[j];
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[j]), "r"(y));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[j + 1]), "r"(y));
				acc[j] = uint32_t(t);
				acc[j + 1] = uint32_t(t >> 32);
			}
		}
		else if (V == 1 || V == 4)
		{
			asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(y));
#pragma unroll
			for (int j = 2; j < NCH; j += 2)
				asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j / 2]), "r"(y));
			asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[NCH]));
			asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[4]), "r"(y));
#pragma unroll
			for (int j = 2; j < NCH; j += 2)
				asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[4 + j / 2]), "r"(y));
			asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[NCH + 1]));
			if (V == 4)
			{
#pragma unroll
				for (int j = 0; j < 8; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(z) : "r"(a[j]));
			}
		}
		else if (V == 2)
		{
			uint64_t t[NCH];
#pragma unroll
			for (int j = 0; j < NCH; ++j) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t[j]) : "r"(a[j]), "r"(y));
			// accumulate the lo words then the hi words, one carry chain each
			asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc[0]) : "r"(uint32_t(t[0])));
#pragma unroll
			for (int j = 1; j < NCH; ++j) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[j]) : "r"(uint32_t(t[j])));
			asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[NCH]));
			asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc[1]) : "r"(uint32_t(t[0] >> 32)));
#pragma unroll
			for (int j = 1; j < NCH; ++j) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[j + 1]) : "r"(uint32_t(t[j] >> 32)));
			asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[NCH + 1]));
		}
		else if (V == 3)
		{
#pragma unroll
			for (int j = 0; j < NCH; ++j) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(y));
		}
		y += 0x1234567u;
	}
	uint32_t r = z;
#pragma unroll
	for (int j = 0; j < NCH + 2; ++j) r ^= acc[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V>
void run(const char* name, double macs_per_iter, uint32_t* d)
{
	int dev_sms = 0;
	cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
	int clk = 0;
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	for (int wps : {1, 2, 4, 8})  // warps per scheduler
	{
		int blocks = dev_sms * wps / 2;  // 256 threads = 8 warps = 2 per scheduler
		if (wps == 1) blocks = dev_sms;
		int threads = wps == 1 ? 128 : 256;
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0);
		cudaEventCreate(&e1);
		k<V><<<blocks, threads>>>(d, 1);
		cudaEventRecord(e0);
		for (int r = 0; r < 5; ++r) k<V><<<blocks, threads>>>(d, r);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		ms /= 5;
		double warps_per_smsp = double(blocks) * threads / 32 / dev_sms / 4;
		double cyc = ms * 1e-3 * clk * 1e3;                                 // cycles (at nominal max clock)
		double warp_macs_per_smsp = warps_per_smsp * ITER * macs_per_iter;  // warp-level MAC instructions per SMSP
		printf("%-34s warps/SMSP=%.0f  %.3f ms  %.2f clk per warp-MAC per SMSP  (%.2e MAC/s)\n", name, warps_per_smsp, ms,
		       cyc / warp_macs_per_smsp, double(blocks) * threads * ITER * macs_per_iter / (ms * 1e-3));
	}
}

int main()
{
	uint32_t* d;
	cudaMalloc(&d, 1 << 24);
	run<0>("A mad.wide (no carry)", NCH, d);
	run<1>("B mad.cc chains (IMAD.WIDE.X)", NCH, d);
	run<2>("C mul.wide + 2 addc", NCH, d);
	run<3>("D mad.lo 32-bit", NCH, d);
	run<4>("E B + 8 independent IADD", NCH, d);
	return 0;
}
