// cg_fused.cuh -- K6f: a whole Jacobi-PCG solve as ONE cooperative kernel (default on one GPU when the rows
// fit on chip; WAVE_CG_FUSED=0 switches it off, =1 also selects it for several ranks).
//
// The three-kernel iteration of kernels.cu (k_spmv, k_cg_update, k_cg_direction) moves
// 12 nnz + 100 n bytes per iteration and pays two launch boundaries.  For problems whose vectors fit
// on chip (about 1.2 M rows on 148 SMs) the iteration can keep x and g in shared memory and h in
// registers for the whole solve, stage the search direction d of the block's column range in shared
// memory once per iteration (the x-gathers of the SpMV then never leave the SM), and separate the
// three phases with grid barriers instead of kernel boundaries:
//     traffic per iteration  12 nnz + 8 n (d written) + ~10 n (d staged) + 8 n (D^-1)
// The arithmetic of one row (ascending columns, separate multiply and add) and of the scalar
// recurrences is that of the three-kernel path and of deal.II's SolverCG (src/WaveNewmark.cpp:256-261).
//
// Measured on B200 (round 2, BASELINE configs[1], Nel=1024 P1, 40 iterations per step): 32 us per CG
// iteration against 48 us of the three-kernel path, 1.38 ms against 2.03 ms per time step; oracle parity
// and parity with the three-kernel path in tests/test_gpu_fused.py.
#pragma once
#include "kernels.cuh"

namespace wv {

constexpr int kFusedThreads = 1024;  // one block per SM: 32 warps, warp w takes slice w of each window
constexpr int kFusedMaxWin = 8;      // windows (kWindow rows) a block keeps on chip

struct CgFusedArgs {
    Sell A;
    const double *val;   // BC-modified system matrix values (SELL layout)
    const double *dinv;  // Jacobi diagonal, row-indexed
    double *x_own;       // solution, owned rows (local-layout vector + own_off)
    double *d;           // search direction, local layout (column-indexed)
    int own_off;
    double *g;           // residual, row-indexed (in: A x - b from the start kernel)
    CgScalars *S;        // in: gh_new/gh_old, reduced_tol, tol, maxit, it = 0, status = 0
    double *partials;    // 2 buffers x gridDim x 2 doubles
    int nwin, wpb;       // windows in the matrix, windows per block
    const int32_t *blk_c0, *blk_cn;  // per block: first staged column and number of staged columns
    int stage_cap;       // doubles of shared memory reserved for the staged part of d
    // several ranks (pc.enabled): the two sums of an iteration also run over the NVLink mailboxes, the
    // boundary blocks of d are stored into the neighbours' ghost blocks in phase 3, flags as in kernels.cu
    PeerComm pc;
    unsigned long long *pub;                 // 8 local words: block 0 publishes the all-rank totals here
    unsigned long long ar_seq0, halo_seq0;   // iteration i uses ar_seq0 + 2i + 1, + 2i + 2 and halo_seq0 + i + 1
};

// per-window column range of the real (unpadded) entries: cmin[w], cmax[w]
void launch_window_col_range(const Launcher &, const Sell &A, int nwin, int32_t *cmin, int32_t *cmax);
// dynamic shared memory the kernel needs for (wpb, stage_cap)
size_t cg_fused_smem_bytes(int wpb, int stage_cap);
// true when the device can co-schedule `grid` blocks of the kernel with that much shared memory
bool cg_fused_supported(int grid, size_t smem_bytes);
// cooperative launch; returns the launch status
cudaError_t launch_cg_fused(const Launcher &, int grid, size_t smem_bytes, const CgFusedArgs &);

}  // namespace wv
