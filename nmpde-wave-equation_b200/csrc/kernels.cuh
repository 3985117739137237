// kernels.cuh -- hand-written sm_100a kernels of the wave-equation hot path.
//
// Everything here is fp64 and HBM-bandwidth bound (O(1) flop/byte): no tensor cores.  Kernel
// ids follow SURVEY.md section 2.1:
//   K1  per-cell quadrature M_e, K_e + CSR scatter          (assemble_matrices)
//   K2  A = M + s K on the shared pattern, Dirichlet rows    (copy_from/add + apply_boundary_values)
//   K3  SELL-32 SpMV (warp per slice) with fused epilogues   (SparseMatrix::vmult)
//   K4  per-cell load vector with the compiled f(x,y,t)      (forcing loops)
//   K5  boundary values -> rhs / start vector                (interpolate_boundary_values)
//   K6  PCG pieces with device-resident scalars, optional    (SolverCG::solve, PreconditionAMG)
//       multigrid V-cycle, NVLink peer all-reduce / halo
//   K7  fused Newmark predictor / corrector + norms          (assemble_rhs z, update_u_v)
//   K8  energy (SpMV + fused dot)                            (compute_and_log_energy)
//   K9  interpolation of u0 / v0 at support points           (VectorTools::interpolate)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "expr_vm.h"
#include "mesh.h"

namespace wv {

// ---- local (per-rank) layout -----------------------------------------------------------------
// A rank owns the DoF blocks of quad rows [jq0, jq1): canonical rows [row0, row0+nown).
// Its vectors cover the canonical range [col0, col0+nloc) = lower ghost block | owned | upper
// ghost block, so a local column index is (global - col0) and halos are contiguous copies.
struct Layout {
    Mesh mesh;
    int jq0, jq1;
    int64_t row0, col0;
    int nown, nloc, own_off;
};

constexpr int kThreads = 256;
constexpr int kMaxRow = 40;

// ---- matrix storage: SELL-32-sigma ---------------------------------------------------------------
// Sliced ELLPACK with C = 32 rows per slice (one warp), rows sorted by descending length inside
// windows of kWindow rows (sigma) so a slice holds rows of one length (P2: 19 / 9; P1: 7) and padding
// is limited to boundary rows.  Slice s stores its entries column-major: element k of the row in
// lane l sits at slice_ptr[s] + 32 k + l, so one warp load of val / col is one contiguous 256 B /
// 128 B segment, and -- because consecutive lanes are consecutive DoFs of one kind -- the k-th
// x gathers of a warp fall on neighbouring addresses.  Entries of a row keep ascending column order
// (the order of the CSR pattern deal.II builds); the CSR row pointer is kept for lengths / export.
// All four value arrays (M, K, SYS1, SYS2) share this one pattern.
constexpr int kSlice = 32;
constexpr int kWindow = 1024;
// ---- translation-invariant rows: the matrix-free ("stencil") operator -------------------------------
// On the structured mesh with a constant wave speed almost every row of M, K and M + sK is a translate
// of one of at most four rows (P1: the vertex row; P2: vertex, horizontal-, vertical-, diagonal-edge
// rows): the same column offsets in storage numbering, in the same (ascending canonical column) order,
// with the same values.  wave_setup detects this numerically (stencil_setup in ctx.cu): a representative
// row per kind is assembled by the same row-gather code, every assembled row is compared with it
// (offsets exactly, values to 1e-12 of the row's largest entry), and a slice whose 32 rows all match one
// kind is marked.  k_spmv then takes the column offsets and values of such a slice from a 4 x 20 table in
// shared memory instead of streaming 12 B per entry from HBM; the other slices (boundary rows, the first
// two DoF lines, rows next to the first quad of a line) and every matrix with variable c keep the SELL
// arrays.  The sum of a row runs over the same entries in the same order with the same arithmetic.
constexpr int kStencilMax = 20;  // entries of the longest translation-invariant row (P2 vertex row: 19)
constexpr int kStencilKinds = 4;
constexpr int kStencilPrefetch = 3;                                  // trips between the L2 prefetch and the use
constexpr int kStencilPad = (kStencilPrefetch + 1) * 148 * 8 * 8;    // >= (prefetch + 1) x warps of the largest grid
struct Sell {
    const uint32_t *slice_ptr;  // nslices + 1, element offsets (multiples of 32)
    const int32_t *col;         // padded local column indices (padding: a valid column, value 0)
    const int32_t *row_of;      // nslices * 32: owned row of (slice, lane), -1 = none
    const int32_t *slot_of;     // nown: slot = slice * 32 + lane of an owned row
    const uint32_t *rowptr;     // CSR row pointer (nown + 1) of the unpadded pattern
    int nslices, nrows;
    int chunk;                  // entries a lane keeps in flight: 7 (P1) or 10 (P2)
    int own_off;                // local column of owned row 0
    // stencil operator (null: none).  slice_info[s] = (kind or -1, first row or -1 when the slice's rows
    // are not consecutive), device memory; st_meta = kStencilKinds x kStencilMax column offsets, then the
    // kStencilKinds row lengths -- a HOST array: the launch copies it into the kernel's parameters
    const int2 *slice_info;
    const int32_t *st_meta;
    const int32_t *sell_list;   // the slices that are not stencil slices, ascending
    int n_sell;
    // the stencil slices in walk order (see k_spmv_st): (slice | ghost << 27 | kind << 28, first row or -1),
    // followed by kStencilPad sentinel entries (-1, -1)
    const int2 *st_walk;
    int n_st;
};

// device-resident CG state (deal.II SolverCG + ReductionControl, src/WaveNewmark.cpp:256-261).
// gh and status are double-buffered by the parity of the iteration index: iteration k reads slot k & 1 and
// its k_cg_direction writes slot (k + 1) & 1, so no kernel ever writes a word that other blocks of the same
// kernel still read, and no kernel needs a last-block tail for the bookkeeping.
struct CgScalars {
    double dAd, gg, gh_new;  // start residual: g.g, g.h (k_spmv result); NCCL-fallback sums; K6f's last values
    double gh[2];            // g.h of the current residual
    double alpha;            // step length of the running iteration (k_cg_update's block 0 -> k_cg_direction)
    double res0, reduced_tol, res;
    double tol, reduce;
    int it, maxit;
    int status[2];           // 0 iterate, 1 success, 2 failure
    int peer_timeout, pad;   // a bounded wait on a peer ran out (reported as WAVE_ERR_CUDA)
};
// where a consumer kernel finds the sums its predecessor produced
enum { SUM_PARTIALS = 0,  // per-block partials of the producer, summed by every consumer block in a fixed order
       SUM_MAILBOX = 1,   // several ranks over NVLink: partials as above + the ranks' totals through the mailboxes
       SUM_SCALAR = 2 };  // totals in CgScalars (NCCL fallback; the multigrid path's g.z)

// ---- NVLink peer exchange inside the CG kernels (nranks > 1) --------------------------------------
// Every rank owns a small mailbox in device memory that all peers map through CUDA IPC.  The two
// per-iteration reductions are one-shot all-reduces split over producer and consumer kernel: the producer
// leaves one partial per block; block 0 of the consumer adds them and stores the rank's total, tagged with
// a sequence number, into every peer's mailbox (P2P stores over NVLink, no waiting); all blocks of the
// consumer collect the other ranks' totals from the local mailbox and add them in rank order, so all ranks
// obtain bitwise identical totals without a separate collective launch and without a producer tail.  The
// halo of the search direction is written by k_cg_direction straight into the neighbours' ghost
// blocks, followed by a flag the next SpMV waits on.  Sequence numbers come from the host and are
// identical on all ranks; waits are bounded (peer_timeout) so a lost peer cannot hang a GPU.
constexpr int kMaxPeers = 8;
struct PeerMailbox {
    // all-reduce slots, "LL" style: every 8-byte word carries 4 bytes of payload and the low 32 bits of
    // the sequence number, so a word validates itself and no fence / flag round trip is needed
    unsigned long long ll[2][kMaxPeers][4];
    unsigned long long halo_flag[2];  // [0] written by the lower neighbour, [1] by the upper one
};
struct PeerComm {
    int enabled, rank, nranks, pad;
    PeerMailbox *box[kMaxPeers];  // box[rank] is the local mailbox
    double *d_lo, *d_hi;          // where my first / last owned block lives in the neighbours' ghost regions
    int lo_count, hi_count;
    int *status;  // &CgScalars::peer_timeout
};

struct SpmvTerm {
    const double *val;
    const double *xa, *xb;  // local-layout vectors; x = ca*xa + cb*xb
    double ca, cb, coef;
    const double *tab;      // HOST array: kStencilKinds x kStencilMax values of the translation-invariant rows (or null)
};
struct SpmvArgs {
    Sell A;
    SpmvTerm t[2];
    const double *add0, *add1;  // row-indexed addends
    double addc0, addc1;
    double *y;                  // row-indexed result (may be null)
    // residual epilogue (CG start): h = dinv*y, d = -h
    const double *dinv;
    double *h_out, *d_out;
    // damped-Jacobi epilogue (multigrid smoother): y_r = jac_x_r + jac_omega * dinv_r * s_r
    const double *jac_x;
    double jac_omega;
    // fused dots: mode 0 none, 1 sum y_i*dotv_i, 2 {sum y_i^2, sum y_i*h_i}
    int dot_mode;
    const double *dotv;
    double *partials;
    unsigned *counter;
    double *result;            // totals (1 or 2 doubles), written by the last block (null with dot_publish)
    // dot_publish (CG iteration): leave the sum to the consumer kernel -- per-block partials, no last-block
    // tail (SUM_PARTIALS / SUM_MAILBOX; with several ranks the consumer also exchanges the ranks' totals)
    int dot_publish;
    const int *skip_flag;      // if non-null and *skip_flag != 0 the kernel returns at once
    int negate;                // set by launch_spmv: coefficient -1 handled as the exact negation of the unit sum
    // peer exchange (see PeerComm): all-reduce of the dot result, wait for the neighbours' halo
    PeerComm pc;
    unsigned long long ar_seq, halo_wait_seq;
    // slices [0, ghost_lo_slices) and [ghost_hi_slice0, nslices) may read ghost entries of x: they wait
    // for the halo flag (lazily, when first reached) and gather through L2 only
    int ghost_lo_slices, ghost_hi_slice0;
};

// ---- launch wrappers (defined in kernels.cu) ---------------------------------------------------
struct Launcher {
    cudaStream_t stream;
    long long *count;    // kernels launched
    cudaError_t *error;  // first launch error seen (sticky; reported at the next synchronisation point)
};

void launch_row_lengths(const Launcher &, const Layout &, uint32_t *rowlen);
// window sort (descending length, stable) -> row_of / slot_of; then per-slice padded sizes
void launch_window_sort(const Launcher &, int nown, int nslots, const uint32_t *rowlen, int32_t *row_of,
                        int32_t *slot_of);
void launch_slice_sizes(const Launcher &, int nslices, const int32_t *row_of, const uint32_t *rowlen,
                        uint32_t *slice_cnt);
void launch_fill_int(const Launcher &, int64_t n, int32_t value, int32_t *dst);
void launch_fill_cols(const Launcher &, const Layout &, const Sell &A, int32_t *col);
void launch_assemble(const Launcher &, const Layout &, const Program *c, const Quadrature *q, const Sell &A,
                     double *M, double *K);
// stencil detection: representative rows (one per kind) -> tables; per-row and per-slice classification.
// counts[0] = rows served by the tables, counts[1] = pattern entries of the other rows
void launch_stencil_tables(const Launcher &, const Layout &, const Program *c, const Quadrature *q, int32_t *st_meta,
                           double *tabM, double *tabK);
void launch_stencil_classify(const Launcher &, const Layout &, const Sell &A, const double *M, const double *K,
                             const int32_t *st_meta, const double *tabM, const double *tabK, int8_t *row_kind,
                             int2 *slice_info, unsigned long long *counts);
void launch_axpy_vals(const Launcher &, int64_t nnz, const double *M, const double *K, double s, double *out);
void launch_find_d0(const Launcher &, const Layout &, const Sell &A, const double *val, double *d0);
void launch_bc_rows(const Launcher &, const Layout &, int nb, const int32_t *brow, const Sell &A, double *val,
                    const double *d0);
void launch_dinv(const Launcher &, const Layout &, const Sell &A, const double *val, int identity, double *dinv);
// SELL -> CSR export of one value array and of the (global) column indices
void launch_canonical_lengths(const Launcher &, const Layout &, const Sell &A, const int32_t *c2i, uint32_t *len);
void launch_export_csr(const Launcher &, const Layout &, const Sell &A, const int32_t *c2i, const int32_t *i2c,
                       const uint32_t *rowptr_c, const double *val, double *csr_val, int32_t *csr_col);
// storage <-> canonical maps of the local range (P2 only) and dst[i] = src[map[i]]
void launch_build_perm(const Launcher &, const Layout &, int32_t *c2i, int32_t *i2c);
void launch_gather(const Launcher &, int n, const int32_t *map, const double *src, double *dst);
void launch_interpolate(const Launcher &, const Layout &, const Program *p, double t, double *vec,
                        double *sx, double *sy);
// load vector in two deterministic passes: per-cell vectors into cellvec (forcing_cells(L) * dofs_per_cell
// doubles), then a row gather in cell order into fvec (owned rows)
int64_t forcing_cells(const Layout &);
void launch_forcing(const Launcher &, const Layout &, const Program *f, const Quadrature *q, double t_np1,
                    double t_n, double w_np1, double w_n, int two_levels, double *cellvec, double *fvec);
// K5 modes
enum { BC_DIRECT = 0, BC_NEWMARK_IMPLICIT = 1, BC_SECOND_DIFF = 2 };
void launch_bc_values(const Launcher &, int mode, int nb, const int32_t *brow, const double *bx,
                      const double *by, const Program *g, double t, double dt, double beta_dt2,
                      const double *z_own, double *x_own, double *rhs, const double *d0);
void launch_spmv(const Launcher &, const SpmvArgs &);
void launch_zero_rows(const Launcher &, int nb, const int32_t *brow, double *vec);
int spmv_grid_blocks(int nslices);
void launch_cg_start(const Launcher &, CgScalars *S);
// blocks a publishing SpMV launch uses (the consumer sums that many partials)
int spmv_launch_blocks(const SpmvArgs &);
int cg_vector_blocks(int n);
// ---- multigrid V-cycle pieces (preconditioner WAVE_PRECOND_MG) -----------------------------------
// x = omega * dinv * b
void launch_scale_rows(const Launcher &, int n, double omega, const double *dinv, const double *b, double *x,
                       const int *skip_flag);
// P1 on mesh Lc (Nel/2) <-> P1 on mesh Lf: x_f += P e_c ; b_c = P^T r_f with Dirichlet rows zeroed.
// Strip-aware: the source vector (ec / rf) is in local layout with valid ghost blocks, the target is the
// owned part of xf (local layout) / bc (row-indexed)
void launch_prolong_add_p1(const Launcher &, const Layout &Lf, const Layout &Lc, const double *ec, double *xf,
                           const int *skip_flag);
void launch_restrict_p1(const Launcher &, const Layout &Lf, const Layout &Lc, const double *rf, double *bc,
                        const int *skip_flag);
// P1 <-> P2 on the same mesh (Lf: r = 2, storage numbering; Lc: r = 1)
void launch_prolong_add_p2p1(const Launcher &, const Layout &Lf, const Layout &Lc, const double *ec, double *xf,
                             const int *skip_flag);
void launch_restrict_p2p1(const Launcher &, const Layout &Lf, const Layout &Lc, const double *rf, double *bc,
                          const int *skip_flag);
// result[0] = g.z ; optionally d = -z (CG start)
void launch_dot_gz(const Launcher &, int n, const double *g, const double *z, double *d_or_null, double *partials,
                   unsigned *counter, double *result, const int *skip_flag);
// One CG iteration k (parity = k & 1) after the SpMV h = A d:
//   k_cg_update:    alpha = gh / dAd ; g += alpha h ; res^2 = g.g ; h = D^-1 g ; gh' = g.h
//   k_cg_direction: x += alpha d ; iteration_status ; beta = gh'/gh ; d = beta d - h (+ halo stores and flag)
// dAd comes from `in` (SUM_* mode, partials of `in_blocks` producer blocks / mailbox sequence in_seq / S->dAd);
// the update's sums go out through `out_partials` (+ mailbox out_seq, or S->gg / S->gh_new).
struct CgSumIo {
    int mode;                  // SUM_*
    const double *partials;    // SUM_PARTIALS: producer's per-block partials
    int blocks, stride;        // SUM_PARTIALS: number of producer blocks, doubles per block
    unsigned long long seq;    // SUM_MAILBOX: sequence number of the exchange
};
void launch_cg_update(const Launcher &, int n, int parity, CgScalars *S, double *g, double *h, const double *dinv,
                      const CgSumIo &in, double *out_partials, unsigned *counter, const PeerComm &pc,
                      unsigned long long out_seq, int out_mode);
// gh_scalar: take g.h' from S->gh_new (multigrid: k_dot_gz wrote it) instead of the second sum of `in`
void launch_cg_direction(const Launcher &, int n, int parity, CgScalars *S, double *x, double *d, const double *h,
                         const CgSumIo &in, int gh_scalar, unsigned *counter, const PeerComm &pc,
                         unsigned long long halo_seq);
void launch_newmark_predict(const Launcher &, int n, double dt, double c1, double c2, double *u, double *v,
                            const double *a);
void launch_newmark_correct(const Launcher &, int n, double cu, double cv, double *u, double *v, const double *a,
                            double *partials, unsigned *counter, double *result);
void launch_norms2(const Launcher &, int n, const double *u, const double *v, double *partials,
                   unsigned *counter, double *result);
void launch_copy(const Launcher &, int n, const double *src, double *dst);
void launch_fill(const Launcher &, int64_t n, double value, double *dst);
void launch_errors(const Launcher &, const Layout &, const Program *sol, const Quadrature *q, double t,
                   const double *u_local, double *partials, unsigned *counter, double *result);
void launch_probe(const Launcher &, const Layout &, double px, double py, const double *u_local, double *out);
void launch_flush_l2(const Launcher &, double *buf, int64_t n);

int reduction_blocks(int n);

}  // namespace wv
