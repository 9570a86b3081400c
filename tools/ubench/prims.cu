// Micro-benchmark of the engine's arithmetic primitives at full occupancy (development aid, `gpurun` only):
// scheduler cycles per warp-level operation = elapsed x SM clock x (SMs x 4 schedulers) / (warps x repetitions).
// One thread per chain, 4 CTAs of 128 threads per SM like the wide kernels; every repetition depends on the previous one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -DQB_HAVE_HORNER_STEP -I qboot_b200/csrc tools/ubench/prims.cu -o tools/ubench/prims
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fastops.cuh"

using namespace qb;
constexpr int L = 32;
using T = mpf<L>;

template <int OP>
__global__ void __launch_bounds__(128, 4) k_prim(const T* a, const T* b, T* out, int reps, int prec)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	T s, c, r;
	copy(s, a[i]);
	copy(c, b[i]);
	for (int k = 0; k < reps; ++k)
	{
		if (OP == 0) fast::mul_small(r, s, int64_t(3 + (k & 255)), prec);
		if (OP == 1) fast::add(r, s, c, prec);               // near operands (same exponent range)
		if (OP == 2) fast::mul(r, s, c, prec);
		if (OP == 3) fast::fma(r, s, c, c, prec);
		if (OP == 4) fast::div(r, s, c, prec);
		if (OP == 5) fast::add_small(r, s, int64_t(1 + (k & 255)), prec);
		if (OP == 6)
		{
			fast::mul_small(r, s, int64_t(3 + (k & 255)), prec);
			fast::add(r, r, c, prec);
		}
#ifdef QB_HAVE_HORNER_STEP
		if (OP == 7)
			if (!fast::horner_step(r, s, uint32_t(3 + (k & 255)), c, prec))
			{
				fast::mul_small(r, s, int64_t(3 + (k & 255)), prec);
				fast::add(r, r, c, prec);
			}
#endif
		// keep magnitudes bounded: fold the exponent back
		if ((r.sk & 3u) == K_REG) r.exp = (r.exp & 7) + 1;
		copy(s, r);
	}
	copy(out[i], s);
}

template <int OP>
double run(const char* name, const T* a, const T* b, T* out, int n, int reps, int prec, double clock_hz, int sms)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k_prim<OP><<<n / 128, 128>>>(a, b, out, 2, prec);
	cudaEventRecord(e0);
	k_prim<OP><<<n / 128, 128>>>(a, b, out, reps, prec);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	double cyc = ms * 1e-3 * clock_hz * sms * 4 / (double(n / 32) * reps);
	printf("%-28s %8.3f ms  %8.1f scheduler-cycles per warp-op\n", name, ms, cyc);
	return cyc;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount, n = sms * 4 * 128 * 2, prec = 1024, reps = 400;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	std::vector<T> ha(n), hb(n);
	srand(1);
	for (int i = 0; i < n; ++i)
	{
		for (int j = 0; j < L; ++j) ha[i].m[j] = uint32_t(rand()) * 2654435761u + j, hb[i].m[j] = uint32_t(rand()) * 40503u + 7 * j;
		ha[i].m[L - 1] |= 0x80000000u;
		hb[i].m[L - 1] |= 0x80000000u;
		ha[i].sk = K_REG;
		hb[i].sk = K_REG | ((i & 1) ? SIGN_BIT : 0u);
		ha[i].exp = 3;
		hb[i].exp = 2 + (i % 5);
	}
	T *a, *b, *o;
	cudaMalloc(&a, sizeof(T) * n);
	cudaMalloc(&b, sizeof(T) * n);
	cudaMalloc(&o, sizeof(T) * n);
	cudaMemcpy(a, ha.data(), sizeof(T) * n, cudaMemcpyHostToDevice);
	cudaMemcpy(b, hb.data(), sizeof(T) * n, cudaMemcpyHostToDevice);
	printf("%s, %d SMs, %d kHz, %d chains, %d repetitions, L = %d\n", p.name, sms, khz, n, reps, L);
	const double hz = khz * 1e3;
	run<0>("mul_small", a, b, o, n, reps, prec, hz, sms);
	run<1>("add (near)", a, b, o, n, reps, prec, hz, sms);
	run<2>("mul", a, b, o, n, reps, prec, hz, sms);
	run<3>("fma", a, b, o, n, reps, prec, hz, sms);
	run<4>("div", a, b, o, n, reps, prec, hz, sms);
	run<5>("add_small", a, b, o, n, reps, prec, hz, sms);
	run<6>("mul_small + add", a, b, o, n, reps, prec, hz, sms);
#ifdef QB_HAVE_HORNER_STEP
	run<7>("horner_step (fused)", a, b, o, n, reps, prec, hz, sms);
#endif
	cudaError_t e = cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
