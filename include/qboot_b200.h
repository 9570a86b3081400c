/* qboot_b200 -- C ABI of the B200 block-table engine.
 *
 * This is the drop-in boundary for qboot's one data-parallel hot path: conformal-block derivative tables
 * gBlock(Delta, spin, S, P) and the F/H blocks built from them.  In the reference the path sits behind the memoised
 * callables of `qboot::Context` (include/qboot/context.hpp:29-56,70-76,123-139, constructed in src/context.cpp:47-61)
 * and the free functions of include/qboot/hor_recursion.hpp:19-30; a maintainer binds these entry points from
 * `Context::gBlock` / `Context::evaluate`.
 *
 * Conventions
 *   - plain C: ints, pointers and sizes only; every call returns 0 or a negative qb_status; nothing throws.
 *   - one handle per GPU; calls on one handle must be serialised by the caller.
 *   - buffers are owned by the caller; the library never frees caller memory.
 *   - a number is a record of W = L + 2 uint32 words, L = ceil(prec / 32):
 *       w[0]  kind | sign << 31   (kind 0 zero, 1 regular, 2 inf, 3 nan)
 *       w[1]  exponent e (int32): value = 0.m * 2^e, the MPFR convention
 *       w[2 .. 2+L)  mantissa limbs, least significant first, left aligned (top bit of w[L+1] set), rounded to prec
 *     i.e. the content of the reference's `mp::real` (include/qboot/mp/real.hpp:87-112) at mp::global_prec = prec.
 *   - all arithmetic is the reference's: each operation rounded once to nearest-even at `prec` bits, performed in the
 *     reference's order, so outputs equal the reference's MPFR results bit for bit.
 *   - there is no CPU fallback: without a CUDA device qb_create fails with QB_ERR_NO_DEVICE.
 */
#ifndef QBOOT_B200_H_
#define QBOOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qb_handle qb_handle;

typedef enum
{
	QB_OK = 0,
	QB_ERR_NO_DEVICE = -1,     /* no CUDA device / CUDA runtime error at creation */
	QB_ERR_BAD_PREC = -2,      /* precision not among the compiled limb counts */
	QB_ERR_NO_CONTEXT = -3,    /* qb_context_init has not been called */
	QB_ERR_BAD_ARG = -4,
	QB_ERR_CUDA = -5,          /* a CUDA call or kernel failed; see qb_last_cuda_error */
	QB_ERR_RANGE = -6,         /* rational / spin out of the supported range */
	QB_ERR_NO_MEMORY = -7,
	QB_ERR_ROUNDING = -8       /* a power stayed next to a rounding boundary at the second Ziv level (probability ~2^-140) */
} qb_status;

/* Version string of the library. */
const char* qb_version(void);
const char* qb_strerror(int status);
/* 1 if kernels for this precision were compiled in (limb count L = ceil(prec/32)), else 0. */
int qb_precision_supported(int prec_bits);

/* Handle = device + stream + scratch arena.  Replaces what `Context`'s two `_task_queue`s own in the reference
 * (include/qboot/context.hpp:29-38).  prec_bits plays the role of mp::global_prec (src/mp/real.cpp:5). */
int qb_create(int device, int prec_bits, qb_handle** out);
void qb_destroy(qb_handle* h);
int qb_words(const qb_handle* h);            /* W = L + 2 */
const char* qb_last_cuda_error(const qb_handle* h);

/* Context(n_Max, lambda, dim) -- src/context.cpp:47-61.  epsilon = (dim - 2) / 2 = eps_num / eps_den exactly.
 * Computes rho = 3 - sqrt(8), the rho->z conversion matrix and the logarithm constants on the host and uploads them. */
int qb_context_init(qb_handle* h, uint32_t n_max, uint32_t lambda, int64_t eps_num, uint32_t eps_den);
/* Read back Context::rho() (W words) and the (lambda+1)^2 row-major rho->z matrix (Context::rho_to_z()). */
int qb_context_get(const qb_handle* h, uint32_t* rho, uint32_t* rho_to_z);
/* dim_M = (lambda+2)^2/4: numbers per table (function_dimension, include/qboot/algebra/complex_function.hpp:43-53) */
int qb_table_dim(const qb_handle* h);
/* numbers per F/H block: sym 0 = Even (H), 1 = Odd (F) */
int qb_block_dim(const qb_handle* h, int sym);

/* gBlock(op, S, P, context) for n operators -- src/hor_recursion.cpp:57-60 (a6 o a4 o a3 o a2 of SURVEY section 8).
 *   spin[n]; delta, S, P: n records each (HOST memory); out: n * dim_M records (HOST memory) in the reference's
 *   flatten() order (dy-major, dx-minor, complex_function.hpp:72-81).
 * Unit operators (Delta = 0) and divergent operators (primary_op.hpp:68-77, averaged as in hor_recursion.cpp:21-28)
 * are handled as the reference handles them.
 * Large batches are cut into chunks (three whole waves of the widest kernel: 11 367 tables at lambda = 19 on a B200);
 * each finished chunk is copied to `out` while the later chunks still
 * compute (pinned `out` makes that copy truly asynchronous; pageable memory works, more slowly).  Returns after the last
 * copy has landed. */
int qb_gblock_batch(qb_handle* h, int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S,
                    const uint32_t* P, uint32_t* out);

/* The same in three steps, so that a caller (or a benchmark) can keep inputs and tables resident in HBM:
 *   plan  : host-side planning (b[0], divergence test, job list) + upload of the inputs
 *   run   : all kernels, asynchronous; chunks of the batch are spread over internal streams that fork from and join
 *           the handle's stream, so work queued on that stream before / after is ordered before / after
 *   fetch : device -> host copy of the n * dim_M records                                                  */
int qb_gblock_plan(qb_handle* h, int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S,
                   const uint32_t* P);
int qb_gblock_run(qb_handle* h);
int qb_gblock_fetch(qb_handle* h, uint32_t* out);
int qb_sync(qb_handle* h);
/* device pointer of the resident tables of the last run (n * dim_M records), for callers that chain kernels */
const uint32_t* qb_gblock_device_tables(const qb_handle* h);
/* number of kernels launched by the library on this handle so far */
int64_t qb_launch_count(const qb_handle* h);
/* CUDA stream of the handle (cudaStream_t as an integer), for event timing on the launching stream */
uint64_t qb_stream(const qb_handle* h);

/* F/H blocks of m terms over the resident tables of the last qb_gblock_run -- Context::F_block / evaluate
 * (include/qboot/context.hpp:123-139): out_term = 2 * proj_sym( v^d * g_table ).
 *   table_idx[m] indexes the resident tables; d: n_d distinct records (HOST); d_idx[m] indexes d; sym[m] in {0, 1};
 *   out: sum over terms of qb_block_dim(sym) records (HOST), terms in order. */
int qb_fblock_batch(qb_handle* h, int m, const int32_t* table_idx, int n_d, const uint32_t* d, const int32_t* d_idx,
                    const int32_t* sym, uint32_t* out);

/* v^d = ((1 - z)(1 - zbar))^d as tables -- ::v_to_d (src/context.cpp:28-42): d = n_d records, out = n_d * dim_M records (HOST). */
int qb_vtod_batch(qb_handle* h, int n_d, const uint32_t* d, uint32_t* out);
/* Context::F_block(d, gBlock, sym) for tables supplied by the caller (include/qboot/context.hpp:123-128): term i multiplies
 * table i (g + i * dim_M records, HOST) by v^d[i]; out as in qb_fblock_batch.  The resident tables are not touched. */
int qb_fblock_tables(qb_handle* h, int m, const uint32_t* g, const uint32_t* d, const int32_t* sym, uint32_t* out);
/* Page-locked host memory for the bulk transfers of this library (tables out, SDPB text out); NULL on failure. */
void* qb_pinned_alloc(size_t bytes);
void qb_pinned_free(void* p);

/* Measurement support.  qb_set_timing(h, 1) brackets each kernel of qb_gblock_run with CUDA events on the handle's
 * stream (chunks then run back to back on that one stream); qb_kernel_times synchronises and returns the device times
 * in ms of the last run: [0] recurrence coefficients (+ their fixed-point copy), [1] series recurrence, [2] powers +
 * rho-derivative chains (scale, Horner), [3] rho->z + transverse recursion.
 * qb_imad_peak measures the limb-product issue rate of this GPU (dependency-free IMAD.WIDE.U32 chains, 32 x 32 + 64 ->
 * 64, at full occupancy) in limb-MACs per second -- the denominator of the roofline for this path. */
int qb_set_timing(qb_handle* h, int on);
int qb_kernel_times(qb_handle* h, float* ms4);
int qb_imad_peak(qb_handle* h, int iters, double* macs_per_s);

/* ---- Multi-GPU: one process per GPU, one handle each; the key list of a job is dealt to the ranks and every rank's tables
 * are gathered in ONE rank's HBM ("one exchange, to rank 0", where the PMP is assembled and written).
 * The exchange is a peer copy over NVLink, not a collective: the root exports its buffer (qb_ipc_export, a 64-byte CUDA IPC
 * handle that travels by any channel -- the benchmark uses torch.distributed), every rank opens it (qb_ipc_open) and names
 * its slice as the sink of its runs; from then on every finished chunk of qb_gblock_run is copied to sink + chunk offset
 * by the copy engine, underneath the kernels of the later chunks, and the handle's stream joins the copies.
 * The sink may equally be pinned host memory or local device memory; NULL switches it off. */
int qb_gblock_set_sink(qb_handle* h, void* dst);
void* qb_device_alloc(qb_handle* h, size_t bytes);   /* cudaMalloc on the handle's device (IPC-exportable); NULL on failure */
void qb_device_free(qb_handle* h, void* p);
int qb_ipc_export(qb_handle* h, const void* dev_ptr, unsigned char* handle64);
int qb_ipc_open(qb_handle* h, const unsigned char* handle64, void** dev_ptr);
int qb_ipc_close(qb_handle* h, void* dev_ptr);

/* Correct rounding of the powers (mpfr_pow / mpfr_ui_pow at src/hor_recursion.cpp:43, real_function.hpp:262, src/context.cpp:31).
 * x^y is evaluated as exp(y ln x) with 96 guard bits; when the discarded bits leave the rounding in doubt (they read
 * 1000...0x or 0111...1x down to bit `bits` of the extended result) the value is recomputed with 192 guard bits (Ziv's
 * strategy).  qb_set_ziv_slack widens or narrows the in-doubt window -- a test knob: with a wide window a large share of
 * all powers goes through the second level; 0 restores the default (56, above the 2^46 error bound of the first level).
 * qb_ziv_counters: out2[0] = powers recomputed at the second level, out2[1] = powers still in doubt there since the handle
 * was created; a new count in out2[1] makes the next qb_gblock_fetch / qb_gblock_batch return QB_ERR_ROUNDING. */
int qb_set_ziv_slack(qb_handle* h, int bits);
int qb_ziv_counters(qb_handle* h, uint64_t* out2);

/* Diagnostic: one arithmetic primitive applied elementwise on the device (n records per operand, HOST memory; unused
 * operands may be NULL).  op: 0 add, 1 sub, 2 mul, 3 div, 4 fma(a*b+c), 6 mul_si(a*ia), 7 add_si, 8 div_si, 9 mul_q
 * (a*ia/ib), 10 div_q, 11 add_q, 12 set_q(ia/ib), 13 rho^b, 14 4^a, 18 si_sub(ia-a), 20 sub_q, 23 cmp(a,b)+1 in word 0,
 * 24 is_integer(a) in word 0.  These are the mpfr_* calls of include/qboot/mp/real.hpp:324-361,767-786 that the table
 * path uses; the entry exists so that the device build of each can be pinned against MPFR.  Needs a context. */
int qb_op_batch(qb_handle* h, int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c,
                const int64_t* ia, const int64_t* ib, uint32_t* out);

/* ---- Polynomial matrix program (PMP) on the device -------------------------------------------------------------------
 * Replaces the host loops of BootstrapEquation::make_cont_mat / make_disc_mat (src/bootstrap_equation.cpp:380-461),
 * PolynomialProgram::create_input (src/polynomial_program.cpp:152-232) and the number formatting of SDPBInput::write
 * (src/sdpb_input.cpp:17-36, include/qboot/mp/real.hpp:987-1030).  The block tables of the last qb_gblock_run stay in
 * HBM; F/H products, the sums over the terms of each crossing equation, the elimination of the normalisation and the
 * conversion to decimal text all happen there, and only the text of the SDPB files crosses PCIe.
 *
 * A *constraint block* is one positivity constraint (sector, operator): sz x sz symmetric matrices at n_samples = D + 1
 * sample points for each of the N = sum(eq_dim) functional components.  Its values are stored as
 *   M[p][n],  p = rc * n_samples + k,  rc = r (r + 1) / 2 + c  (c <= r, the lower triangle),  n = 0 .. N - 1
 * which is the row order of the reference's d_B (rows enumerate (r, c <= r, k), polynomial_program.cpp:205-221).        */
typedef struct qb_pmp_term
{
	int32_t eq;   /* index of the equation the term belongs to */
	int32_t rc;   /* matrix position: max(r,c) (max(r,c) + 1) / 2 + min(r,c) */
	int32_t tab0; /* block table of sample k = tab[tab0 + k], an index into the tables of the last qb_gblock_run */
	int32_t d;    /* index into `d` (the power of v: (Delta_2 + Delta_3) / 2 of the term's externals) */
	int32_t coef; /* index into `coef`: term.coeff() / (r == c ? 1 : 2) already rounded as the reference rounds it; -1 = 1 */
} qb_pmp_term;
typedef struct qb_pmp_block
{
	int32_t n_samples, sz; /* D + 1 and the matrix size */
	int32_t term0, n_terms; /* this block's terms, in the order the equations list them */
} qb_pmp_block;
/* Assemble all blocks of a program: for every block, position, sample and component
 *   M = sum over the terms of that (equation, position), in order, of coef * [2 proj_sym(v^d g_table)]_component
 * added onto +0 (bootstrap_equation.cpp:399-405,447-453).  eq_dim[e] must equal qb_block_dim(eq_sym[e]).
 * d: n_d records, coef: n_coef records (HOST).  Drops the blocks of any earlier program on this handle. */
int qb_pmp_assemble(qb_handle* h, int n_eq, const int32_t* eq_dim, const int32_t* eq_sym, int n_blocks,
                    const qb_pmp_block* blocks, int n_terms, const qb_pmp_term* terms, int n_tab, const int32_t* tab,
                    int n_d, const uint32_t* d, int n_coef, const uint32_t* coef);
/* Add a block computed elsewhere (discrete sectors contracted with OPE vectors, user-supplied inequalities):
 * mat = schur * N records in the M[p][n] layout, target = schur records or NULL for zero.  Returns its block index. */
int qb_pmp_block_add(qb_handle* h, int n_samples, int sz, const uint32_t* mat, const uint32_t* target);
/* Read a block back: schur * N records. */
int qb_pmp_block_fetch(qb_handle* h, int block, uint32_t* out);
/* PolynomialProgram::create_input for the blocks sel[0..n_sel): with the n_elim variables lead[e] eliminated through
 *   y[lead_e] + sum_m eqn[e][free_m] y[free_m] = eqn_target[e]
 * d_B[p][m] = -M[p][free_m] and d_c[p] = -target[p], then for e = 0 .. n_elim - 1 an fma with M[p][lead_e]
 * (polynomial_program.cpp:205-221).  eqn: n_elim * N records, eqn_target: n_elim records (HOST). */
int qb_pmp_pack(qb_handle* h, int n_sel, const int32_t* sel, int n_elim, const int32_t* lead, int n_free,
                const int32_t* free_idx, const uint32_t* eqn, const uint32_t* eqn_target);
/* d_B (schur * n_free records, row-major) and d_c (schur records) of the j-th selected block; either may be NULL. */
int qb_pmp_packed_fetch(qb_handle* h, int j, uint32_t* B, uint32_t* c);
/* SDPB text of every packed number, formatted on the device: `ndigits` significant digits, mpfr_get_str rounding, the
 * reference's default-float layout, one number per line.  sizes[2 j] / sizes[2 j + 1] = bytes of the d_B / d_c text of
 * selected block j; qb_pmp_text_fetch copies the concatenation (block 0 d_B, block 0 d_c, block 1 d_B, ...) to HOST.
 * Returns 0, or the count of numbers whose exponent is outside the device formatter's range: read them with
 * qb_pmp_text_declined (idx = position in the stream, recs = the records), format them on the host and hand the text
 * back with qb_pmp_text_patch (text i at texts + i * stride, len[i] bytes incl. the newline), which completes the call. */
int qb_pmp_text(qb_handle* h, int ndigits, int64_t* sizes);
int qb_pmp_text_declined(qb_handle* h, int cap, int64_t* idx, uint32_t* recs);
int qb_pmp_text_patch(qb_handle* h, int cnt, const int64_t* idx, const char* texts, int stride, const int32_t* len,
                      int64_t* sizes);
int qb_pmp_text_fetch(qb_handle* h, char* out);
/* out[0] blocks in the arena, [1] N, [2] records in the arena, [3] selected blocks of the last pack */
int qb_pmp_info(const qb_handle* h, int64_t* out4);
/* Bilinear bases at the sample points -- the other half of PolynomialProgram::create_input (src/polynomial_program.cpp:224-229):
 *   out[i] = RN( eval(poly[items[i].poly], x[items[i].x]) * s[items[i].s] )
 * with eval = Horner's rule by fused multiply-adds, exactly Polynomial::eval with a real argument
 * (include/qboot/algebra/polynomial.hpp:66-83).  Polynomial p has the coefficients coef[poly_ofs[p] .. poly_ofs[p + 1]) (constant
 * term first; none = the zero polynomial).  Everything HOST memory in, n_items records out.  Needs a context. */
typedef struct qb_eval_item
{
	int32_t poly, x, s;
} qb_eval_item;
int qb_poly_eval_batch(qb_handle* h, int n_polys, const int32_t* poly_ofs, const uint32_t* coef, int n_x, const uint32_t* x, int n_s,
                       const uint32_t* s, int n_items, const qb_eval_item* items, uint32_t* out);
/* Diagnostic: n numbers (HOST) formatted on the device into slots of `slot` >= ndigits + 16 bytes; len[i] = bytes used
 * or a negative code when declined (see format_decimal in qboot_b200/csrc/pmp.cuh). */
int qb_dec_format(qb_handle* h, int ndigits, int n, const uint32_t* x, char* slots, int slot, int32_t* len);

#ifdef __cplusplus
}
#endif

#endif /* QBOOT_B200_H_ */
