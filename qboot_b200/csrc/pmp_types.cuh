// qboot_b200 -- plain-data descriptors of the PMP stage (shared by the launch declarations, the kernels and the engine;
// kept apart from pmp.cuh so that editing the PMP device code does not rebuild the block-table kernels).
#pragma once

#include <cstdint>

#include "mpf.cuh"

namespace qb
{
	// a term of a crossing equation inside one constraint block (host builds these, see include/qboot_b200.h)
	struct PmpTerm
	{
		int32_t eq;    // equation index
		int32_t rc;    // packed lower-triangle position r (r + 1) / 2 + c, c <= r
		int32_t tab0;  // table of sample k: tab[tab0 + k]
		int32_t d;     // index of the v-power table
		int32_t coef;  // index of the coefficient record (already divided by 2 off the diagonal), < 0: coefficient 1
	};
	struct PmpBlock
	{
		int32_t n_samples, sz, term0, n_terms;
		int64_t ofs;      // first record of M in the arena; M[p][n], p = rc * n_samples + k
		int64_t thread0;  // first thread of this block in the assembly grid (n_rc * N * n_samples threads)
		int64_t pack0;    // first thread of this block in the pack grid (schur * (n_free + 1) threads)
		int64_t ofs_b;    // first record of d_B in its arena (schur x n_free, row-major)
		int64_t ofs_c;    // first record of d_c in its arena (schur)
		int64_t ofs_t;    // first record of the target values (schur), < 0: zero target
	};
	QB_HD QB_INLINE int pmp_schur(const PmpBlock& b) { return b.n_samples * (b.sz * (b.sz + 1) / 2); }

	// one entry of qb_poly_eval_batch: polynomial, argument and scale by index
	struct EvalItem
	{
		int32_t poly, x, s;
	};

	// Tables for format_decimal, built on the host for one (L, ndigits): powers of five in two levels
	//   5^k = P13[k / 13] * 5^(k % 13),  P13[a] = 5^(13 a) as little-endian limbs (length plen[a], offset pofs[a])
	// and the bounds 10^ndigits, 10^(ndigits-1) as L + 2 limbs.
	constexpr int QB_DEC_MAXA = 64;  // 5^(13 * 63): covers k <= 831
	struct DecTables
	{
		const uint32_t* p13;       // concatenated limbs of 5^(13 a)
		int32_t pofs[QB_DEC_MAXA + 1];
		int ndigits;
		const uint32_t* ten_n;     // 10^ndigits,     L + 2 limbs
		const uint32_t* ten_n1;    // 10^(ndigits-1), L + 2 limbs
	};
}  // namespace qb
