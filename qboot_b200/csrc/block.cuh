// qboot_b200 -- from the rho-series to the block-derivative table and the F/H blocks (host + device building blocks).
//
//   deriv_*          : (1/k!) d^k/drho^k [ (4 rho)^Delta h(rho) ] at rho*, k = 0..lambda
//                      (reference: gBlock_real, src/hor_recursion.cpp:40-56; RealFunctionWithPower::derivate /
//                      approximate, include/qboot/algebra/real_function.hpp:236-264)
//   rho_to_z         : rho-Taylor -> z-Taylor coefficients (RealConverter::convert, real_function.hpp:208-213,
//                      dot(Vector, Matrix), include/qboot/algebra/matrix.hpp:444-456)
//   expand_off_diag  : transverse derivatives by the Casimir recursion (Context::expand_off_diagonal, src/context.cpp:62-133)
//   v_to_d           : table of ((1-z)(1-zbar))^d (src/context.cpp:28-42)
//   fblock_entry     : one entry of 2 * proj_sym(v^d * g) (Context::F_block, include/qboot/context.hpp:123-128;
//                      ComplexFunction mul / proj, include/qboot/algebra/complex_function.hpp:190-206)
//
// Every function performs the reference's operations in the reference's order, each rounded at `prec` bits.
// Operand placement as in hor.cuh: anywhere; the functions say where their results go.
#pragma once

#include "hor.cuh"
#include "mpf_math.cuh"

namespace qb
{
	// packed index of (dx, dy) in a Mixed table of order lambda: dy-major, dx-minor (complex_function.hpp:72-81)
	QB_HD QB_INLINE int cf_index(int lambda, int dx, int dy) { return (lambda + 2 - dy) * dy + dx; }
	QB_HD QB_INLINE int cf_dim(int lambda) { return (lambda + 2) * (lambda + 2) / 4; }
	// packed index inside one parity class (sym 0 = Even, 1 = Odd)
	QB_HD QB_INLINE int cf_index_sym(int lambda, int sym, int dx, int dy)
	{
		return sym == 0 ? (lambda / 2 * 2 + 3 - dy) * dy / 2 + dx / 2 : ((lambda - 1) / 2 * 2 + 3 - dy) * dy / 2 + (dx - 1) / 2;
	}
	QB_HD QB_INLINE int cf_dim_sym(int lambda, int sym)
	{
		int h = sym == 0 ? (lambda + 2) / 2 : (lambda + 1) / 2;
		return h * (h + 1) / 2;
	}

	// ---- rho-derivatives (all operands thread-local) -----------------------------------------------------------------
	// level-k scaling of one coefficient: a <- a_prev * (q + n)      (derivate, real_function.hpp:236-240)
	template <int L>
	QB_HD QB_INLINE void deriv_scale(mpf<L>& a, const mpf<L>& a_prev, const mpf<L>& q, uint32_t n, int prec)
	{
		mpf<L> f;
		fast::add_small(f, q, int64_t(n), prec);
		fast::mul(a, a_prev, f, prec);
	}
	// Horner step s <- s * rho + a                                   (approximate, real_function.hpp:248-259)
	template <int L>
	QB_HD QB_INLINE void deriv_horner(mpf<L>& s, const mpf<L>& rho, const mpf<L>& a, int prec)
	{
		fast::fma(s, s, rho, a, prec);
	}
	// f_k = s * rho^q / k!                                           (real_function.hpp:262, hor_recursion.cpp:46-53)
	template <int L>
	QB_HD QB_INLINE void deriv_finish(mpf<L>& f, const mpf<L>& s, const mpf<L>& rho_pow_q, const mpf<L>& kfact, uint32_t k,
	                                  int prec)
	{
		mul(f, s, rho_pow_q, prec);
		if (k >= 1) div(f, f, kfact, prec);
	}

	// z = x[0] * M[0][i], then += x[j] * M[j][i]   (dot, matrix.hpp:444-456); x, M anywhere; M row-major (lambda+1)^2;
	// z thread-local
	template <int L>
	QB_HD QB_NOINLINE void rho_to_z_entry(mpf<L>& z, const mpf<L>* x, int64_t xstride, const mpf<L>* M, int lambda, int i,
	                                      int prec)
	{
		mpf<L> t, xv, mv;
		copy(xv, x[0]);
		copy(mv, M[i]);
		fast::mul(z, xv, mv, prec);
		for (int j = 1; j <= lambda; ++j)
		{
			copy(xv, x[int64_t(j) * xstride]);
			copy(mv, M[j * (lambda + 1) + i]);
			if (is_zero(mv) && kind(xv) == K_REG)
			{
				// The matrix is triangular and the reference multiplies through its zeros: x * 0 = +-0 exactly, and adding
				// a zero leaves a nonzero z unchanged; zero + zero is -0 only if both are.  Taken inline so that the lanes
				// of a warp (different columns, different zero patterns) stay converged.
				if (kind(z) == K_REG) continue;
				if (is_zero(z))
				{
					z.sk &= ~SIGN_BIT | ((xv.sk ^ mv.sk) & SIGN_BIT);
					continue;
				}
			}
			fast::mul(t, xv, mv, prec);
			fast::add(z, z, t, prec);
		}
	}

	// ---- transverse recursion ------------------------------------------------------------------------------------------
	// quad_casimir4 = 4 * ((2 eps + spin) spin / 2 + (Delta / 2 - eps - 1) Delta)    (primary_op.hpp:60-63, context.cpp:70)
	// r, D thread-local
	template <int L>
	QB_HD QB_NOINLINE void quad_casimir4(mpf<L>& r, const mpf<L>& D, uint32_t spin, Rational eps, int prec)
	{
		mpf<L> t;
		div_si(t, D, 2, prec);
		sub_q(t, t, eps.num, eps.den, prec);
		add_si(t, t, -1, prec);
		mul(t, t, D, prec);
		// rational part (2 eps + spin) * spin / 2 = (2 num + spin den) spin / (2 den)
		int64_t rn = (2 * eps.num + int64_t(spin) * int64_t(eps.den)) * int64_t(spin);
		uint32_t rd = 2u * eps.den;
		add_q(t, t, rn, rd, prec);
		mul_si(r, t, 4, prec);
	}

	// One row entry of the recursion: everything except the same-row correction, i.e. val / common_factor.
	// f is the Mixed table (packed, anywhere), rows n-1 and n-2 must be complete; val, S, P, qc4 thread-local.
	// (context.cpp:72-121)
	//
	// The 36 operations of the entry are run by a small interpreter -- a step counter decoded into (op, dst, src, src2 /
	// integer) and ONE instance of every primitive -- instead of straight-line code: with all arithmetic force-inlined
	// the straight-line version was 63 000 instructions (1 MB) per kernel and ran at the speed of instruction fetch
	// (ncu: 5-10 no-instruction stall cycles per issue).  Operation order and operands are exactly the reference's.
	// LOCKSTEP (device): the CTA meets at a barrier in front of every step, so that its warps run the same primitive at
	// the same time and one instruction fetch serves all of them (threads without work pass active = false and only
	// keep the barriers company).
	template <int L, bool LOCKSTEP = false>
	QB_HD QB_NOINLINE void offdiag_val(mpf<L>& val, const mpf<L>* f, int64_t fs, int lambda, int m, int n, const mpf<L>& S,
	                                   const mpf<L>& P, const mpf<L>& qc4, Rational eps, int prec, bool active = true)
	{
		enum : int { VAL = 0, TERM = 1, T2 = 2, H = 3, RS = 4, RP = 5, RQ = 6 };
		enum : int { LOADF, MULS, ADDS, ADD, MUL, SETQ, ADDQ, DIVQ, NOP };
		mpf<L> R[7];
		copy(R[RS], S);
		copy(R[RP], P);
		copy(R[RQ], qc4);
		const int64_t M = m, N = n, en = eps.num;
		QB_ROLL
		for (int step = 0; step < 41; ++step)
		{
			int op = NOP, d = 0, a = 0, b = 0;
			int64_t v = 0;
#if defined(__CUDA_ARCH__)
			if (LOCKSTEP) __syncthreads();
#endif
			switch (active ? step : -1)
			{
			// h[m + 2, n - 1]
			case 0: op = LOADF; v = cf_index(lambda, m + 2, n - 1); break;
			case 1: op = MULS; d = VAL; a = H; v = -(M + 1) * (M + 2); break;
			// h[m + 1, n - 1]
			case 2: op = ADDQ; d = TERM; a = RS; v = en; break;
			case 3: op = MULS; d = TERM; a = TERM; v = 2; break;
			case 4: op = ADDS; d = TERM; a = TERM; v = -(M + 4 * N - 6); break;
			case 5: op = MULS; d = TERM; a = TERM; v = 2 * (M + 1); break;
			case 6: op = LOADF; v = cf_index(lambda, m + 1, n - 1); break;
			case 7: op = MUL; d = TERM; a = TERM; b = H; break;
			case 8: op = ADD; d = VAL; a = VAL; b = TERM; break;
			// h[m, n - 1]
			case 9: op = SETQ; d = TERM; v = en * (4 * (M + N - 1)); break;
			case 10: op = ADDS; d = TERM; a = TERM; v = M * (M + 8 * N - 5) + N * (4 * N - 2) - 2; break;
			case 11: op = MULS; d = T2; a = RP; v = 2; break;
			case 12: op = ADD; d = TERM; a = TERM; b = T2; break;
			case 13: op = MULS; d = T2; a = RS; v = 8 * N + 4 * M - 8; break;
			case 14: op = ADD; d = TERM; a = TERM; b = T2; break;
			case 15: op = ADD; d = TERM; a = TERM; b = RQ; break;
			case 16: op = MULS; d = TERM; a = TERM; v = 4; break;
			case 17: op = LOADF; v = cf_index(lambda, m, n - 1); break;
			case 18: op = MUL; d = TERM; a = TERM; b = H; break;
			case 19: op = ADD; d = VAL; a = VAL; b = TERM; break;
			// h[m - 1, n - 1]
			case 20: if (m >= 1) { op = SETQ; d = TERM; v = en * (2 * (M + 1 - 2 * N)); } break;
			case 21: if (m >= 1) { op = ADDS; d = TERM; a = TERM; v = M * (M + 12 * N - 13) + N * (12 * N - 34) + 22; } break;
			case 22: if (m >= 1) { op = MULS; d = T2; a = RP; v = 2; } break;
			case 23: if (m >= 1) { op = ADD; d = TERM; a = TERM; b = T2; } break;
			case 24: if (m >= 1) { op = MULS; d = T2; a = RS; v = 8 * N + 2 * M - 10; } break;
			case 25: if (m >= 1) { op = ADD; d = TERM; a = TERM; b = T2; } break;
			case 26: if (m >= 1) { op = MULS; d = TERM; a = TERM; v = 8; } break;
			case 27: if (m >= 1) { op = LOADF; v = cf_index(lambda, m - 1, n - 1); } break;
			case 28: if (m >= 1) { op = MUL; d = TERM; a = TERM; b = H; } break;
			case 29: if (m >= 1) { op = ADD; d = VAL; a = VAL; b = TERM; } break;
			// h[m + 1, n - 2] and h[m + 2, n - 2]
			case 30: if (n >= 2) { op = SETQ; d = TERM; v = -2 * en; } break;
			case 31: if (n >= 2) { op = ADDS; d = TERM; a = TERM; v = 3 * M + 4 * N - 6; } break;
			case 32: if (n >= 2) { op = MULS; d = T2; a = RS; v = 2; } break;
			case 33: if (n >= 2) { op = ADD; d = TERM; a = TERM; b = T2; } break;
			case 34: if (n >= 2) { op = MULS; d = TERM; a = TERM; v = 8 * (M + 1); } break;
			case 35: if (n >= 2) { op = LOADF; v = cf_index(lambda, m + 1, n - 2); } break;
			case 36: if (n >= 2) { op = MUL; d = TERM; a = TERM; b = H; } break;
			case 37: if (n >= 2) { op = LOADF; v = cf_index(lambda, m + 2, n - 2); } break;
			case 38: if (n >= 2) { op = MULS; d = T2; a = H; v = 4 * (M + 1) * (M + 2); } break;
			case 39: if (n >= 2) { op = ADD; d = TERM; a = TERM; b = T2; } break;
			case 40: if (n >= 2) { op = ADD; d = VAL; a = VAL; b = TERM; } break;
			default: break;
			}
			switch (op)
			{
			case LOADF: copy(R[H], f[v * fs]); break;
			case MULS: fast::mul_small(R[d], R[a], v, prec); break;
			case ADDS: fast::add_small(R[d], R[a], v, prec); break;
			case ADD: fast::add(R[d], R[a], R[b], prec); break;
			case MUL: fast::mul(R[d], R[a], R[b], prec); break;
			case SETQ: set_q(R[d], v, eps.den, prec); break;
			case ADDQ: add_q(R[d], R[a], v, eps.den, prec); break;
			default: break;
			}
		}
		// common_factor = 4 n eps + (4 n - 2) n
		int64_t cn = int64_t(4 * n) * eps.num + int64_t(4 * n - 2) * int64_t(n) * int64_t(eps.den);
		if (active) div_q(val, R[VAL], cn, eps.den, prec);
	}
	// h[m, n] = val + 8 h[m-3, n] + 4 h[m-2, n] - 2 h[m-1, n]; the three predecessors are handed in thread-local
	// (p1 = h[m-1,n], p2 = h[m-2,n], p3 = h[m-3,n]); result thread-local                          (context.cpp:123-130)
	template <int L>
	QB_HD QB_NOINLINE void offdiag_close(mpf<L>& out, int m, const mpf<L>& val, const mpf<L>& p1, const mpf<L>& p2,
	                                     const mpf<L>& p3, int prec)
	{
		mpf<L> term, t2;
		set_zero(term);
		if (m >= 3)
		{
			fast::mul_small(t2, p3, 8, prec);
			fast::add(term, term, t2, prec);
		}
		if (m >= 2)
		{
			fast::mul_small(t2, p2, 4, prec);
			fast::add(term, term, t2, prec);
		}
		if (m >= 1)
		{
			fast::mul_small(t2, p1, 2, prec);
			fast::add(term, term, t2, prec, true);
		}
		fast::add(out, val, term, prec);
	}
	// whole table, serial: f (anywhere) row 0 (dx = 0..lambda) must be filled; D, S, P anywhere
	template <int L>
	QB_HD void expand_off_diag(mpf<L>* f, int64_t fs, int lambda, const mpf<L>& Din, uint32_t spin, const mpf<L>& Sin,
	                           const mpf<L>& Pin, Rational eps, int prec)
	{
		mpf<L> D, S, P, qc4, val, p1, p2, p3, cur;
		copy(D, Din);
		copy(S, Sin);
		copy(P, Pin);
		quad_casimir4(qc4, D, spin, eps, prec);
		for (int n = 1; n <= lambda / 2; ++n)
		{
			set_zero(p1);
			set_zero(p2);
			set_zero(p3);
			for (int m = 0; m + 2 * n <= lambda; ++m)
			{
				offdiag_val(val, f, fs, lambda, m, n, S, P, qc4, eps, prec);
				offdiag_close(cur, m, val, p1, p2, p3, prec);
				copy(f[int64_t(cf_index(lambda, m, n)) * fs], cur);
				copy(p3, p2);
				copy(p2, p1);
				copy(p1, cur);
			}
		}
	}

	// ---- v^d table -------------------------------------------------------------------------------------------------------
	// f(0,0) = 4^-d given (pow); f(0,n) = f(0,n-1) * 4 * (-d + (n-1)) / n; f(m,n) = f(m-1,n) * (-4 d + 2 (m+2n-1)) / m
	// f anywhere; d, four_pow_minus_d anywhere
	template <int L>
	QB_HD void v_to_d_table(mpf<L>* f, int64_t fs, int lambda, const mpf<L>& din, const mpf<L>& four_pow_minus_d, int prec)
	{
		mpf<L> d, md, t, u, row0, cur;
		copy(d, din);
		copy(md, d);
		neg(md);
		copy(row0, four_pow_minus_d);
		for (int n = 0; n <= lambda / 2; ++n)
		{
			if (n > 0)
			{
				mul_si(t, row0, 4, prec);
				add_si(u, md, int64_t(n - 1), prec);
				mul(t, t, u, prec);
				div_si(row0, t, int64_t(n), prec);
			}
			copy(f[int64_t(cf_index(lambda, 0, n)) * fs], row0);
			copy(cur, row0);
			for (int m = 1; m + 2 * n <= lambda; ++m)
			{
				mul_si(u, d, -4, prec);
				add_si(u, u, int64_t(2 * (m + 2 * n - 1)), prec);
				mul(t, cur, u, prec);
				div_si(cur, t, int64_t(m), prec);
				copy(f[int64_t(cf_index(lambda, m, n)) * fs], cur);
			}
		}
	}

	// ---- F / H block -------------------------------------------------------------------------------------------------------
	// out = 2 * sum_{dy1 <= N} sum_{dx1 <= M} v(dx1, dy1) * g(M - dx1, N - dy1), accumulated in that order.
	// v, g anywhere; out thread-local
	template <int L>
	QB_HD QB_NOINLINE void fblock_entry(mpf<L>& out, const mpf<L>* v, int64_t vs, const mpf<L>* g, int64_t gs, int lambda, int M,
	                                    int N, int prec)
	{
		mpf<L> acc, t, a, b;
		set_zero(acc);
		for (int dy1 = 0; dy1 <= N; ++dy1)
			for (int dx1 = 0; dx1 <= M; ++dx1)
			{
				copy(a, v[int64_t(cf_index(lambda, dx1, dy1)) * vs]);
				copy(b, g[int64_t(cf_index(lambda, M - dx1, N - dy1)) * gs]);
				fast::mul(t, a, b, prec);
				fast::add(acc, acc, t, prec);
			}
		mul_si(out, acc, 2, prec);
	}
}  // namespace qb
