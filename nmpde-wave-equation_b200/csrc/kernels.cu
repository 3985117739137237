// kernels.cu -- sm_100a kernels of the wave-equation hot path (see kernels.cuh for the map to the
// reference's call sites).  All arithmetic fp64, all kernels HBM-bound; grids are sized in
// multiples of the SM count where the work is a grid-stride stream.
#include "kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cub/block/block_radix_sort.cuh>

namespace wv {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSMs = 148;  // B200

// ---- deterministic reductions -------------------------------------------------------------------
// Block tree reduction of NV values (fixed order), result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
    __shared__ double sm[NV * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(kFull, v[k], off);
    __syncthreads();  // protect sm against a previous use
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) sm[warp * NV + k] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = lane < nwarps ? sm[lane * NV + k] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(kFull, x, off);
            v[k] = x;
        }
    }
}

// Grid-wide sum: every block contributes its NV partial sums; the block that arrives last adds the
// per-block partials in block order and returns true with the totals in thread 0's v[].
// Fixed grid => bitwise reproducible.  The counter is left at zero for the next launch.
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double *partials, unsigned *counter) {
    __shared__ bool s_last;
    block_sum<NV>(v);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[(size_t)blockIdx.x * NV + k] = v[k];
        __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
            s += __ldcg(&partials[(size_t)b * NV + k]);
        v[k] = s;
    }
    block_sum<NV>(v);
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}

// Last-block detection without a reduction (for scalar bookkeeping after all blocks have read it).
// sys = true when the blocks wrote to peer memory: the fence then has system scope.
__device__ __forceinline__ bool last_block_done(unsigned *counter, bool sys = false) {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sys) __threadfence_system();
        else __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        s_last = (ticket == gridDim.x - 1);
        if (s_last) *counter = 0u;
    }
    __syncthreads();
    return s_last;
}

// ---- peer exchange over NVLink (see PeerComm in kernels.cuh) -------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *p, unsigned long long seq) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < seq) {
        if (clock64() - t0 > 6000000000LL) return false;  // ~3 s: give up instead of hanging the GPU
    }
    return true;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// One-shot all-reduce of NV (<= 2) doubles in two halves.  p2p_publish (first warp of one block per rank):
// lane r stores this rank's totals into rank r's mailbox (one NVLink hop, no ordering needed, no waiting).
// p2p_collect (first warp of any block): lane r polls rank r's words in the local mailbox; the sum runs in
// rank order on every rank (bitwise identical everywhere).  Every 8-byte word carries 4 bytes of payload
// and the low 32 bits of the sequence number, so a word validates itself.  The producer publishes in its
// tail; the consumer kernel collects in its first instructions, so the NVLink flight time and the ranks'
// skew overlap with the consumer's launch instead of extending the producer.
template <int NV>
__device__ __forceinline__ void p2p_publish(const PeerComm &pc, unsigned long long seq, const double (&vals)[NV]) {
    const int lane = threadIdx.x & 31;
    const int slot = (int)(seq & 1ull);
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __shfl_sync(kFull, vals[k], 0);
    if (lane < pc.nranks) {
        unsigned long long *dst = pc.box[lane]->ll[slot][pc.rank];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v[k]);
            st_relaxed_sys(&dst[2 * k], (bits & 0xffffffffull) | tag);
            st_relaxed_sys(&dst[2 * k + 1], (bits >> 32) | tag);
        }
    }
}
// `self` (may be null): this rank's own totals, taken from registers instead of the mailbox
template <int NV>
__device__ __forceinline__ void p2p_collect(const PeerComm &pc, unsigned long long seq, double (&out)[NV],
                                            const double *self = nullptr) {
    const int lane = threadIdx.x & 31;
    const int slot = (int)(seq & 1ull);
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    bool ok = true;
    double got[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) got[k] = 0.0;
    if (self && lane == pc.rank) {
#pragma unroll
        for (int k = 0; k < NV; ++k) got[k] = self[k];
    } else if (lane < pc.nranks) {
        const unsigned long long *src = pc.box[pc.rank]->ll[slot][lane];
        const long long t0 = clock64();
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            unsigned long long lo, hi;
            for (;;) {
                lo = ld_relaxed_sys(&src[2 * k]);
                hi = ld_relaxed_sys(&src[2 * k + 1]);
                if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                if (clock64() - t0 > 6000000000LL) { ok = false; break; }  // ~3 s: give up, do not hang
                if (blockIdx.x != 0) __nanosleep(100);  // hundreds of blocks poll one line: leave it to block 0
            }
            got[k] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
        }
    }
    ok = __all_sync(kFull, ok);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (int r = 0; r < pc.nranks; ++r) t += __shfl_sync(kFull, got[k], r);  // rank order on every rank
        out[k] = ok ? t : __longlong_as_double(0x7ff8000000000000LL);  // NaN stops the solve (status 2)
    }
    if (!ok && lane == 0) *pc.status = 1;
}

// ---- sums handed from a producer kernel to its consumer (see SUM_* in kernels.cuh) --------------------
// Producer tail.  SUM_PARTIALS and SUM_MAILBOX: one partial per block and nothing else (no fence, no
// ticket, no second pass) -- the consumer finishes the sum.  SUM_SCALAR: the last block adds the partials
// in block order and writes the totals to `scalars` (an NCCL all-reduce follows).
template <int NV>
__device__ __forceinline__ void produce_sums(double (&v)[NV], int mode, double *partials, unsigned *counter,
                                             double *scalars) {
    if (mode != SUM_SCALAR) {
        block_sum<NV>(v);
        if (threadIdx.x == 0)
#pragma unroll
            for (int k = 0; k < NV; ++k) partials[(size_t)blockIdx.x * NV + k] = v[k];
        return;
    }
    if (grid_sum<NV>(v, partials, counter) && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) scalars[k] = v[k];
    }
}
// Consumer head: the totals in every thread of every block, bitwise identical in all of them.  Every block
// adds the producer's per-block partials in the same fixed order.  With several ranks (SUM_MAILBOX) block 0
// then stores this rank's total into every peer's mailbox and all blocks collect the other ranks' totals
// from the local mailbox and add them in rank order (their own from registers): the exchange starts the
// moment the producer grid has drained -- one NVLink flight on the critical path, no producer tail.
template <int NV>
__device__ __forceinline__ void consume_sums(const CgSumIo &in, const PeerComm &pc, const double *scalars,
                                             double (&out)[NV]) {
    __shared__ double s_bc[NV];
    if (in.mode == SUM_SCALAR) {
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = scalars[k];
        return;
    }
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = 0.0;
    for (int b = threadIdx.x; b < in.blocks; b += blockDim.x)
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] += __ldcg(&in.partials[(size_t)b * in.stride + k]);
    block_sum<NV>(v);  // totals of this rank in thread 0
    if (in.mode == SUM_MAILBOX) {
        if (threadIdx.x < 32) {
            double loc[NV], tot[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) loc[k] = __shfl_sync(kFull, v[k], 0);
            if (blockIdx.x == 0) p2p_publish<NV>(pc, in.seq, loc);
            p2p_collect<NV>(pc, in.seq, tot, loc);
            if (threadIdx.x == 0)
#pragma unroll
                for (int k = 0; k < NV; ++k) s_bc[k] = tot[k];
        }
    } else if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) s_bc[k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = s_bc[k];
}

// ---- sparsity pattern: DoFTools::make_sparsity_pattern + compress (src/WaveNewmark.cpp:33-35) ----
// One thread per entity slot; the row is the sorted union of the DoFs of the adjacent cells.
// cols: canonical ids in ascending order (the CSR order); icols: the same entries' storage ids
__device__ int build_row(const Mesh &m, int i, int j, int kind, int64_t *cols, int64_t *icols) {
    int64_t cells[6];
    const int nc = entity_cells(m, i, j, kind, cells);
    const int dpc = dofs_per_cell(m.r);
    int n = 0;
    for (int c = 0; c < nc; ++c) {
        int64_t d[6], di[6];
        cell_dofs(m, cells[c], d);
        cell_dofs_internal(m, cells[c], di);
        for (int k = 0; k < dpc; ++k) {
            // sorted insert, skip duplicates
            const int64_t v = d[k];
            int p = n;
            while (p > 0 && cols[p - 1] > v) --p;
            if (p > 0 && cols[p - 1] == v) continue;
            for (int q = n; q > p; --q) { cols[q] = cols[q - 1]; icols[q] = icols[q - 1]; }
            cols[p] = v;
            icols[p] = di[k];
            ++n;
        }
    }
    return n;
}

struct SlotRange {
    int jlo, jhi, nk;  // lattice rows [jlo, jhi], kinds per lattice point
};
__host__ __device__ inline SlotRange owned_slots(const Layout &L) {
    return {L.jq0, L.jq1, L.mesh.r == 1 ? 1 : 4};
}
__host__ __device__ inline SlotRange local_slots(const Layout &L) {
    const int jlo = L.jq0 > 0 ? L.jq0 - 1 : 0;
    const int jhi = L.jq1 + 1 <= L.mesh.ny ? L.jq1 + 1 : L.mesh.ny;
    return {jlo, jhi, L.mesh.r == 1 ? 1 : 4};
}
__host__ __device__ inline int64_t slot_count(const Layout &L, const SlotRange &s) {
    return (int64_t)(s.jhi - s.jlo + 1) * (L.mesh.nx + 1) * s.nk;
}
__device__ __forceinline__ void decode_slot(const Layout &L, const SlotRange &s, int64_t t, int &i, int &j,
                                            int &kind) {
    kind = (int)(t % s.nk);
    const int64_t q = t / s.nk;
    i = (int)(q % (L.mesh.nx + 1));
    j = s.jlo + (int)(q / (L.mesh.nx + 1));
}

__global__ void k_row_lengths(Layout L, uint32_t *rowlen) {
    const SlotRange s = owned_slots(L);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, t, i, j, kind);
    const int64_t dof = entity_dof_internal(L.mesh, i, j, kind);
    if (dof < L.row0 || dof >= L.row0 + L.nown) return;
    int64_t cols[kMaxRow], icols[kMaxRow];
    rowlen[dof - L.row0] = (uint32_t)build_row(L.mesh, i, j, kind, cols, icols);
}

// ---- SELL-32-sigma construction --------------------------------------------------------------------
// One block per window of kWindow rows: stable sort by descending row length (radix sort of the key
// (maxlen - len) << 16 | local index); slot = position after the sort.
__global__ void __launch_bounds__(kWindow) k_window_sort(int nown, const uint32_t *rowlen, int32_t *row_of,
                                                         int32_t *slot_of) {
    using Sort = cub::BlockRadixSort<unsigned, kWindow, 1>;
    __shared__ typename Sort::TempStorage tmp;
    const int wbase = blockIdx.x * kWindow;
    const int row = wbase + threadIdx.x;
    const unsigned len = row < nown ? rowlen[row] : 0u;
    unsigned key[1] = {((unsigned)(kMaxRow - len) << 16) | (unsigned)threadIdx.x};
    Sort(tmp).Sort(key, 0, 24);
    const int src = wbase + (int)(key[0] & 0xffffu);
    const int slot = wbase + threadIdx.x;
    if (src < nown) {
        row_of[slot] = src;
        slot_of[src] = slot;
    } else
        row_of[slot] = -1;
}
// padded size of a slice = 32 * (longest row in it); one warp per slice
__global__ void k_slice_sizes(int nslices, const int32_t *row_of, const uint32_t *rowlen, uint32_t *slice_cnt) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = row_of[s * kSlice + lane];
    unsigned len = r >= 0 ? rowlen[r] : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) len = max(len, __shfl_xor_sync(kFull, len, off));
    if (lane == 0) slice_cnt[s] = len * kSlice;
}
__global__ void k_fill_int(int64_t n, int32_t value, int32_t *dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

__device__ __forceinline__ uint32_t sell_row_base(const Sell &A, int row) {
    const int slot = A.slot_of[row];
    return A.slice_ptr[slot >> 5] + (uint32_t)(slot & 31);
}
__device__ __forceinline__ uint32_t sell_row_len(const Sell &A, int row) { return A.rowptr[row + 1] - A.rowptr[row]; }

__global__ void k_fill_cols(Layout L, Sell A, int32_t *col) {
    const SlotRange s = owned_slots(L);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, t, i, j, kind);
    const int64_t dof = entity_dof_internal(L.mesh, i, j, kind);
    if (dof < L.row0 || dof >= L.row0 + L.nown) return;
    int64_t cols[kMaxRow], icols[kMaxRow];
    const int n = build_row(L.mesh, i, j, kind, cols, icols);
    const uint32_t base = sell_row_base(A, (int)(dof - L.row0));
    for (int k = 0; k < n; ++k) col[base + kSlice * k] = (int32_t)(icols[k] - L.col0);
}

// position of (row, local storage column c) in the padded arrays; rows are short (<= 19 entries) and
// ordered by canonical column, so the storage ids are searched linearly
__device__ __forceinline__ uint32_t find_col(const Sell &A, int row, int32_t c) {
    const uint32_t base = sell_row_base(A, row), len = sell_row_len(A, row);
    for (uint32_t k = 0; k < len; ++k)
        if (A.col[base + kSlice * k] == c) return base + kSlice * k;
    return 0xffffffffu;
}

// canonical row lengths (for the exported CSR row pointer), then the export itself: canonical row r is
// storage row c2i[r]; its entries are already in ascending canonical column order
__global__ void k_canonical_lengths(Layout L, Sell A, const int32_t *c2i, uint32_t *len) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= L.nown) return;
    const int row = (c2i ? c2i[r + L.own_off] : r + L.own_off) - L.own_off;
    len[r] = sell_row_len(A, row);
}
__global__ void k_export_csr(Layout L, Sell A, const int32_t *c2i, const int32_t *i2c, const uint32_t *rowptr_c,
                             const double *val, double *csr_val, int32_t *csr_col) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= L.nown) return;
    const int row = (c2i ? c2i[r + L.own_off] : r + L.own_off) - L.own_off;
    const uint32_t base = sell_row_base(A, row), len = sell_row_len(A, row), o = rowptr_c[r];
    for (uint32_t k = 0; k < len; ++k) {
        if (csr_val) csr_val[o + k] = val[base + kSlice * k];
        if (csr_col) {
            const int32_t ic = A.col[base + kSlice * k];
            csr_col[o + k] = (int32_t)((i2c ? i2c[ic] : ic) + L.col0);
        }
    }
}
// storage <-> canonical permutation of the local range (identity for P1)
__global__ void k_build_perm(Layout L, int32_t *c2i, int32_t *i2c) {
    const SlotRange s = local_slots(L);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, tid, i, j, kind);
    const int64_t c = entity_dof(L.mesh, i, j, kind);
    if (c < L.col0 || c >= L.col0 + L.nloc) return;
    const int64_t t = entity_dof_internal(L.mesh, i, j, kind);
    c2i[c - L.col0] = (int32_t)(t - L.col0);
    i2c[t - L.col0] = (int32_t)(c - L.col0);
}
// dst[i] = src[map[i]]
__global__ void __launch_bounds__(kThreads) k_gather(int n, const int32_t *__restrict__ map,
                                                     const double *__restrict__ src, double *__restrict__ dst) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[map[i]];
}

// ---- K1: assemble_matrices (src/WaveNewmark.cpp:56-108 == src/WaveTheta.cpp:56-108) ---------------
// Row gather, no atomics: one thread per owned DoF row.  For each adjacent cell, in ascending cell
// order (entity_cells) -- the order in which the reference's serial cell loop adds into the global
// entry -- the thread integrates its own row of the element matrices with the reference's quadrature
// loop (q outer, j inner; c evaluated at the quadrature point at parser time 0) and adds it to the
// row's entries.  Every entry is therefore a fixed sequence of additions: matrices are bitwise
// reproducible between contexts and identical for any number of ranks.
template <int R>
__device__ void assemble_row(const Mesh &m, int i, int j, int kind, const Program *cprog, const Quadrature &Q,
                             int n, const int64_t *icols, double *Mrow, double *Krow) {
    constexpr int DPC = R == 1 ? 3 : 6;
    int64_t cells[6];
    const int nc = entity_cells(m, i, j, kind, cells);
    const int64_t self = entity_dof_internal(m, i, j, kind);
    for (int k = 0; k < n; ++k) { Mrow[k] = 0.0; Krow[k] = 0.0; }
    for (int c = 0; c < nc; ++c) {
        int64_t d[6];
        cell_dofs_internal(m, cells[c], d);
        int a = 0;
#pragma unroll
        for (int k = 0; k < DPC; ++k)
            if (d[k] == self) a = k;
        double X0, Y0, sx, sy;
        cell_geometry(m, cells[c], X0, Y0, sx, sy);
        const double det = sx * sy, adet = fabs(det);
        const double ix = sy / det, iy = sx / det;  // J^{-T} = diag(1/sx, 1/sy) written as the adjugate / det
        double Me[DPC], Ke[DPC];
#pragma unroll
        for (int b = 0; b < DPC; ++b) { Me[b] = 0.0; Ke[b] = 0.0; }
        for (int q = 0; q < Q.nq; ++q) {
            double phi[DPC], dxi[DPC], deta[DPC], gx[DPC], gy[DPC];
            shape_values(R, Q.xi[q], Q.eta[q], phi);
            shape_grads(R, Q.xi[q], Q.eta[q], dxi, deta);
#pragma unroll
            for (int b = 0; b < DPC; ++b) { gx[b] = ix * dxi[b]; gy[b] = iy * deta[b]; }
            double pa = phi[0], gxa = gx[0], gya = gy[0];
#pragma unroll
            for (int b = 1; b < DPC; ++b)
                if (a == b) { pa = phi[b]; gxa = gx[b]; gya = gy[b]; }
            const double JxW = Q.w[q] * adet;
            const double xq = X0 + sx * Q.xi[q], yq = Y0 + sy * Q.eta[q];
            const double cv = eval(cprog, xq, yq, 0.0);
            const double c2 = cv * cv;
#pragma unroll
            for (int b = 0; b < DPC; ++b) {
                Me[b] += pa * phi[b] * JxW;
                Ke[b] += c2 * (gxa * gx[b] + gya * gy[b]) * JxW;
            }
        }
#pragma unroll
        for (int b = 0; b < DPC; ++b) {
            int p = 0;
            while (p < n - 1 && icols[p] != d[b]) ++p;
            Mrow[p] += Me[b];
            Krow[p] += Ke[b];
        }
    }
}
template <int R>
__global__ void __launch_bounds__(128) k_assemble(Layout L, const Program *cprog, Quadrature Q, Sell A,
                                                  double *M, double *K) {
    const SlotRange s = owned_slots(L);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, t, i, j, kind);
    const int64_t dof = entity_dof_internal(L.mesh, i, j, kind);
    if (dof < L.row0 || dof >= L.row0 + L.nown) return;
    int64_t cols[kMaxRow], icols[kMaxRow];
    double Mrow[kMaxRow], Krow[kMaxRow];
    const int n = build_row(L.mesh, i, j, kind, cols, icols);
    assemble_row<R>(L.mesh, i, j, kind, cprog, Q, n, icols, Mrow, Krow);
    const uint32_t base = sell_row_base(A, (int)(dof - L.row0));
    for (int k = 0; k < n; ++k) {
        M[base + kSlice * k] = Mrow[k];
        K[base + kSlice * k] = Krow[k];
    }
}

// ---- stencil detection (see kernels.cuh) -------------------------------------------------------------
// Representative rows: the entities of the mesh's middle quad, assembled by the row-gather code above.
// The choice does not depend on the rank, so every rank holds bitwise identical tables.
template <int R>
__global__ void k_stencil_tables(Layout L, const Program *cprog, Quadrature Q, int32_t *st_meta, double *tabM,
                                 double *tabK) {
    const int kind = threadIdx.x;
    const int nk = R == 1 ? 1 : kStencilKinds;
    if (blockIdx.x != 0 || kind >= kStencilKinds) return;
    if (kind >= nk) { st_meta[kStencilKinds * kStencilMax + kind] = 0; return; }
    const int i = L.mesh.nx / 2, j = L.mesh.ny / 2;
    int64_t cols[kMaxRow], icols[kMaxRow];
    double Mrow[kMaxRow], Krow[kMaxRow];
    const int n = build_row(L.mesh, i, j, kind, cols, icols);
    assemble_row<R>(L.mesh, i, j, kind, cprog, Q, n, icols, Mrow, Krow);
    const int64_t self = entity_dof_internal(L.mesh, i, j, kind);
    st_meta[kStencilKinds * kStencilMax + kind] = n <= kStencilMax ? n : 0;
    for (int k = 0; k < kStencilMax; ++k) {
        const bool in = k < n && n <= kStencilMax;
        st_meta[kind * kStencilMax + k] = in ? (int32_t)(icols[k] - self) : 0;
        tabM[kind * kStencilMax + k] = in ? Mrow[k] : 0.0;
        tabK[kind * kStencilMax + k] = in ? Krow[k] : 0.0;
    }
}
// one thread per owned row: does it equal its kind's representative (offsets exactly, values to 1e-12)?
__global__ void k_classify_rows(Layout L, Sell A, const double *__restrict__ M, const double *__restrict__ K,
                                const int32_t *__restrict__ st_meta, const double *__restrict__ tabM,
                                const double *__restrict__ tabK, int8_t *row_kind) {
    const SlotRange s = owned_slots(L);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, t, i, j, kind);
    const int64_t dof = entity_dof_internal(L.mesh, i, j, kind);
    if (dof < L.row0 || dof >= L.row0 + L.nown) return;
    const int row = (int)(dof - L.row0);
    const int len = st_meta[kStencilKinds * kStencilMax + kind];
    bool ok = len > 0 && !entity_on_boundary(L.mesh, i, j, kind) && (int)sell_row_len(A, row) == len;
    if (ok) {
        const uint32_t base = sell_row_base(A, row);
        double mM = 0.0, mK = 0.0;
        for (int k = 0; k < len; ++k) {
            mM = fmax(mM, fabs(tabM[kind * kStencilMax + k]));
            mK = fmax(mK, fabs(tabK[kind * kStencilMax + k]));
        }
        for (int k = 0; k < len && ok; ++k) {
            const uint32_t q = base + kSlice * k;
            ok = A.col[q] - (row + L.own_off) == st_meta[kind * kStencilMax + k] &&
                 fabs(M[q] - tabM[kind * kStencilMax + k]) <= 1e-12 * mM &&
                 fabs(K[q] - tabK[kind * kStencilMax + k]) <= 1e-12 * mK;
        }
    }
    row_kind[row] = ok ? (int8_t)kind : (int8_t)-1;
}
// one warp per slice: all 32 rows present and of one matching kind -> stencil slice
__global__ void k_classify_slices(Sell A, const int8_t *__restrict__ row_kind, int2 *slice_info,
                                  unsigned long long *counts) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= A.nslices) return;
    const int r = A.row_of[s * kSlice + lane];
    const int k = r >= 0 ? (int)row_kind[r] : -1;
    const int k0 = __shfl_sync(kFull, k, 0), r0 = __shfl_sync(kFull, r, 0);
    const bool same = __all_sync(kFull, k == k0 && k >= 0);
    const bool consecutive = __all_sync(kFull, r == r0 + lane);
    unsigned nnz = (!same && r >= 0) ? sell_row_len(A, r) : 0u;
    unsigned rows = same ? 1u : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        nnz += __shfl_down_sync(kFull, nnz, off);
        rows += __shfl_down_sync(kFull, rows, off);
    }
    if (lane == 0) {
        slice_info[s] = make_int2(same ? k0 : -1, consecutive ? r0 : -1);
        if (rows) atomicAdd(&counts[0], (unsigned long long)rows);
        if (nnz) atomicAdd(&counts[1], (unsigned long long)nnz);
    }
}

// ---- K2: matrix_a = M + s K on the shared pattern (src/WaveNewmark.cpp:111-112) --------------------
__global__ void __launch_bounds__(kThreads) k_axpy_vals(int64_t nnz, const double *__restrict__ M,
                                                        const double *__restrict__ K, double s,
                                                        double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride)
        out[e] = M[e] + s * K[e];
}

// [deal.II] MatrixTools::apply_boundary_values (Trilinos overload, Release build):
// d0 = |first non-zero diagonal entry| of the rank's rows; each boundary row becomes d0 * e_i.
__global__ void k_find_d0(Layout L, Sell A, const double *val, double *d0) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double r = 0.0;
    for (int i = 0; i < L.nown; ++i) {
        const uint32_t pos = find_col(A, i, i + L.own_off);
        if (pos != 0xffffffffu && val[pos] != 0.0) { r = fabs(val[pos]); break; }
    }
    *d0 = r;
}
__global__ void k_bc_rows(Layout L, int nb, const int32_t *brow, Sell A, double *val, const double *d0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int row = brow[b];
    const double diag = *d0;
    const uint32_t base = sell_row_base(A, row), len = sell_row_len(A, row);
    for (uint32_t k = 0; k < len; ++k)
        val[base + kSlice * k] = (A.col[base + kSlice * k] == row + L.own_off) ? diag : 0.0;
}
// Jacobi: 1 / diag(A) (stand-in for PreconditionAMG / PreconditionSSOR per the north star)
__global__ void k_dinv(Layout L, Sell A, const double *val, int identity, double *dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nown) return;
    if (identity) { dinv[i] = 1.0; return; }
    dinv[i] = 1.0 / val[find_col(A, i, i + L.own_off)];
}

// ---- K9: VectorTools::interpolate at DoF support points (src/WaveNewmark.cpp:292-293) --------------
__global__ void k_interpolate(Layout L, const Program *p, double t, double *vec, double *sx, double *sy) {
    const SlotRange s = local_slots(L);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, tid, i, j, kind);
    const int64_t dof = entity_dof(L.mesh, i, j, kind);
    if (dof < L.col0 || dof >= L.col0 + L.nloc) return;
    double x, y;
    entity_point(L.mesh, i, j, kind, x, y);
    if (vec) vec[entity_dof_internal(L.mesh, i, j, kind) - L.col0] = eval(p, x, y, t);  // storage order
    if (sx) { sx[dof - L.col0] = x; sy[dof - L.col0] = y; }                             // canonical order
}

// ---- K4: forcing load vector (src/WaveNewmark.cpp:151-171, src/WaveTheta.cpp:151-180) ---------------
// Two passes, no atomics: (1) one thread per cell integrates the cell's load vector into cellvec
// (cell-major, DPC entries per cell); (2) one thread per owned row adds the entries of its adjacent
// cells in ascending cell order, the order of the reference's serial cell loop.
template <int R>
__global__ void __launch_bounds__(128) k_forcing_cells(Layout L, const Program *f, Quadrature Q, double t_np1,
                                                       double t_n, double w_np1, double w_n, int two_levels,
                                                       double *cellvec) {
    constexpr int DPC = R == 1 ? 3 : 6;
    const int jtop = L.jq1 < L.mesh.ny ? L.jq1 : L.mesh.ny - 1;
    const int jbot = L.jq0;  // block j holds the DoFs first met in quad row j: owned rows touch quad rows jq0 .. jq1
    const int64_t c_begin = 2LL * jbot * L.mesh.nx, c_end = 2LL * (jtop + 1) * L.mesh.nx;
    const int64_t cell = c_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= c_end) return;
    double X0, Y0, sx, sy;
    cell_geometry(L.mesh, cell, X0, Y0, sx, sy);
    const double adet = fabs(sx * sy);
    double acc[DPC];
#pragma unroll
    for (int a = 0; a < DPC; ++a) acc[a] = 0.0;
    for (int q = 0; q < Q.nq; ++q) {
        double phi[DPC];
        shape_values(R, Q.xi[q], Q.eta[q], phi);
        const double JxW = Q.w[q] * adet;
        const double xq = X0 + sx * Q.xi[q], yq = Y0 + sy * Q.eta[q];
        double fv;
        if (two_levels) {
            const double f_n = eval(f, xq, yq, t_n);
            const double f_np1 = eval(f, xq, yq, t_np1);
            fv = w_np1 * f_np1 + w_n * f_n;
        } else
            fv = eval(f, xq, yq, t_np1);
#pragma unroll
        for (int a = 0; a < DPC; ++a) acc[a] += fv * phi[a] * JxW;
    }
#pragma unroll
    for (int a = 0; a < DPC; ++a) cellvec[(cell - c_begin) * DPC + a] = acc[a];
}
template <int R>
__global__ void __launch_bounds__(128) k_forcing_gather(Layout L, const double *__restrict__ cellvec,
                                                        double *__restrict__ fvec) {
    constexpr int DPC = R == 1 ? 3 : 6;
    const SlotRange s = owned_slots(L);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(L, s)) return;
    int i, j, kind;
    decode_slot(L, s, t, i, j, kind);
    const int64_t dof = entity_dof_internal(L.mesh, i, j, kind);
    if (dof < L.row0 || dof >= L.row0 + L.nown) return;
    const int64_t c_begin = 2LL * L.jq0 * L.mesh.nx;
    int64_t cells[6];
    const int nc = entity_cells(L.mesh, i, j, kind, cells);
    double sum = 0.0;
    for (int c = 0; c < nc; ++c) {
        int64_t d[6];
        cell_dofs_internal(L.mesh, cells[c], d);
        int a = 0;
#pragma unroll
        for (int k = 0; k < DPC; ++k)
            if (d[k] == dof) a = k;
        sum += cellvec[(cells[c] - c_begin) * DPC + a];
    }
    fvec[dof - L.row0] = sum;
}

// ---- K5: boundary values (src/WaveNewmark.cpp:186-241, :348-374; src/WaveTheta.cpp:259-272) ---------
__global__ void k_bc_values(int mode, int nb, const int32_t *brow, const double *bx, const double *by,
                            const Program *g, double t, double dt, double beta_dt2, const double *z_own,
                            double *x_own, double *rhs, const double *d0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int row = brow[b];
    const double x = bx[b], y = by[b];
    double val;
    if (mode == BC_DIRECT)
        val = eval(g, x, y, t);
    else if (mode == BC_NEWMARK_IMPLICIT)
        val = (eval(g, x, y, t) - z_own[row]) / beta_dt2;
    else {
        const double inv_dt2 = 1.0 / (dt * dt);
        val = (eval(g, x, y, t) - 2.0 * eval(g, x, y, t - dt) + eval(g, x, y, t - 2.0 * dt)) * inv_dt2;
    }
    x_own[row] = val;
    rhs[row] = val * (*d0);
}

// ---- K3: SELL-32 SpMV (TrilinosWrappers::SparseMatrix::vmult) ---------------------------------------
// One warp per slice, one row per lane.  Every warp load of val / col is one contiguous segment; the
// loop is unrolled so 4 x (col, val) loads are in flight per lane before the dependent x gathers.
// A row is accumulated in ascending column order with separate multiply and add roundings, i.e. the
// arithmetic of a serial CSR loop (bitwise equal to it for a single term).  Epilogues: addends, the
// CG residual start (h = D^-1 g, d = -h) and fused dot products (deterministic grid reduction).
// streaming loads for the matrix arrays (read once per launch): evict-first, keep L1/L2 for x
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

// row epilogue shared by both SpMV kernels: addends, damped-Jacobi update, CG residual start, fused dots
__device__ __forceinline__ void spmv_epilogue(const SpmvArgs &a, int r, double s, double (&dots)[2]) {
    if (a.negate) s = -s;
    if (a.add0) s += a.addc0 * a.add0[r];
    if (a.add1) s += a.addc1 * a.add1[r];
    if (a.jac_x) s = a.jac_x[r] + a.jac_omega * (a.dinv[r] * s);
    if (a.y) a.y[r] = s;
    double hv = 0.0;
    if (a.h_out) {
        hv = a.dinv[r] * s;
        a.h_out[r] = hv;
        a.d_out[r] = -hv;
    }
    if (a.dot_mode == 1) dots[0] += s * a.dotv[r];
    else if (a.dot_mode == 2) { dots[0] += s * s; dots[1] += s * hv; }
}
__device__ __forceinline__ void spmv_tail(const SpmvArgs &a, double (&dots)[2]) {
    if (!a.dot_mode) return;
    if (a.dot_publish)  // CG iteration: the consumer kernel (k_cg_update) finishes the sum
        produce_sums<2>(dots, SUM_PARTIALS, a.partials, a.counter, nullptr);
    else if (grid_sum<2>(dots, a.partials, a.counter) && threadIdx.x == 0) {
        a.result[0] = dots[0];
        if (a.dot_mode == 2) a.result[1] = dots[1];
    }
}
// wait (once per warp, lazily) for the neighbours' halo of this iteration's search direction
__device__ __forceinline__ void halo_wait(const SpmvArgs &a, int lane) {
    if (lane == 0) {
        bool ok = true;
        PeerMailbox *box = a.pc.box[a.pc.rank];
        if (a.pc.rank > 0) ok = spin_until(&box->halo_flag[0], a.halo_wait_seq) && ok;
        if (a.pc.rank < a.pc.nranks - 1) ok = spin_until(&box->halo_flag[1], a.halo_wait_seq) && ok;
        if (!ok) *a.pc.status = 1;
    }
    __syncwarp();
}

template <int NT, bool TWOX, int CH, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_spmv(SpmvArgs a) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (a.skip_flag && *a.skip_flag != 0) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (kThreads / 32);
    double dots[2] = {0.0, 0.0};
    // Persistent warps, grid-stride over slices.  Per chunk a lane issues CH col loads and CH*NT val
    // loads back to back (unconditional: a short tail re-reads the row's last entry and is masked out
    // of the sum), then the CH dependent x gathers, then accumulates in column order.
    // The walk starts in the middle of the matrix (rotation by nslices/2): the few slices that read
    // ghost entries sit at both ends of the row range and are reached mid-kernel, when the neighbours'
    // halo (written by their k_cg_direction over NVLink) has long arrived -- the wait below is off
    // the critical path.  Those slices gather through L2 (ld.global.cg): an L1 line that straddles
    // the owned / ghost boundary may have been filled before the halo landed.
    const int rot = a.halo_wait_seq ? a.A.nslices / 2 : 0;
    bool halo_ready = a.halo_wait_seq == 0;
    int it = (blockIdx.x * kThreads + threadIdx.x) >> 5;  // position in the walk
    int slice = it + rot < a.A.nslices ? it + rot : it + rot - a.A.nslices;
    uint32_t b = 0, e = 0;
    int r = -1;
    if (it < a.A.nslices) {
        b = a.A.slice_ptr[slice];
        e = a.A.slice_ptr[slice + 1];
        r = a.A.row_of[slice * kSlice + lane];
    }
    while (it < a.A.nslices) {
        // metadata of this warp's next slice, requested before the long-latency work below
        const int nit = it + nwarps;
        const int nslice = nit + rot < a.A.nslices ? nit + rot : nit + rot - a.A.nslices;
        uint32_t nb = 0, ne = 0;
        int nr = -1;
        if (nit < a.A.nslices) {
            nb = a.A.slice_ptr[nslice];
            ne = a.A.slice_ptr[nslice + 1];
            nr = a.A.row_of[nslice * kSlice + lane];
        }
        const bool ghosty = a.pc.enabled && (slice < a.ghost_lo_slices || slice >= a.ghost_hi_slice0);
        if (ghosty && !halo_ready) {  // warp-uniform
            halo_wait(a, lane);
            halo_ready = true;
        }
        const int len = (int)((e - b) >> 5);
        const uint32_t base = b + lane;
        double s = 0.0;
        for (int k0 = 0; k0 < len; k0 += CH) {
            int c[CH];
            double v[NT][CH], x[NT][CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int kk = min(k0 + k, len - 1);
                const uint32_t q = base + (uint32_t)kk * kSlice;
                c[k] = ld_stream(&a.A.col[q]);
#pragma unroll
                for (int t = 0; t < NT; ++t) v[t][k] = ld_stream(&a.t[t].val[q]);
            }
#pragma unroll
            for (int k = 0; k < CH; ++k)
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const double xa = ghosty ? __ldcg(&a.t[t].xa[c[k]]) : a.t[t].xa[c[k]];
                    double xv = __dmul_rn(a.t[t].ca, xa);
                    if (TWOX && a.t[t].xb)
                        xv = __dadd_rn(xv, __dmul_rn(a.t[t].cb, ghosty ? __ldcg(&a.t[t].xb[c[k]]) : a.t[t].xb[c[k]]));
                    x[t][k] = xv;
                }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (k0 + k < len) {  // warp-uniform
                    double prod = __dmul_rn(a.t[0].coef, __dmul_rn(v[0][k], x[0][k]));
#pragma unroll
                    for (int t = 1; t < NT; ++t)
                        prod = __dadd_rn(prod, __dmul_rn(a.t[t].coef, __dmul_rn(v[t][k], x[t][k])));
                    s = __dadd_rn(s, prod);
                }
            }
        }
        if (r >= 0) spmv_epilogue(a, r, s, dots);
        it = nit; slice = nslice; b = nb; e = ne; r = nr;
    }
    spmv_tail(a, dots);
}

// ---- K3s: the same operator with the translation-invariant rows served from tables (kernels.cuh) ------
// Phase 1 walks all slices and takes the stencil ones.  The column offsets and values of the (at most
// four) representative rows travel as a kernel parameter, i.e. in the constant bank: inside the
// `switch (kind)` every row is a fully unrolled sequence with compile-time length (P1: 7; P2: 19 for a
// vertex row, 9 for the three edge kinds) whose offsets and values are constant-bank operands of the
// index add and of the multiply.  A row costs its x gathers (L1 / L2: every x entry is used by ~11 rows),
// 8 B of y and the epilogue's vectors -- no matrix stream, ~5 instructions per entry.  All gathers of a
// row are in flight before the sum, which runs over the same entries in the same order with the same
// separate multiply / add roundings as the SELL path (multiplications by a coefficient of exactly 1 are
// skipped: they are exact).  Phase 2 takes the few remaining slices (boundary rows, first DoF lines) from
// the SELL arrays through a precomputed list, with short chunks.
struct StencilParams {
    int off[kStencilKinds][kStencilMax];
    double val[2][kStencilKinds][kStencilMax];
};
template <int LEN, int BATCH, int NT, bool TWOX, bool UNIT, bool CG>
__device__ __forceinline__ double stencil_row(const SpmvArgs &a, const StencilParams &p, const int kind, int cself) {
    double s = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < LEN; k0 += BATCH) {
        double x[NT][BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            if (k0 + k < LEN) {
                const int idx = cself + p.off[kind][k0 + k];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    double xv = CG ? __ldcg(&a.t[t].xa[idx]) : a.t[t].xa[idx];
                    if (!UNIT) {
                        xv = __dmul_rn(a.t[t].ca, xv);
                        if (TWOX && a.t[t].xb)
                            xv = __dadd_rn(xv, __dmul_rn(a.t[t].cb, CG ? __ldcg(&a.t[t].xb[idx]) : a.t[t].xb[idx]));
                    }
                    x[t][k] = xv;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            if (k0 + k < LEN) {
                double prod = __dmul_rn(p.val[0][kind][k0 + k], x[0][k]);
                if (!UNIT) {
                    prod = __dmul_rn(a.t[0].coef, prod);
#pragma unroll
                    for (int t = 1; t < NT; ++t)
                        prod = __dadd_rn(prod, __dmul_rn(a.t[t].coef, __dmul_rn(p.val[t][kind][k0 + k], x[t][k])));
                }
                s = __dadd_rn(s, prod);
            }
        }
    }
    return s;
}
template <bool P2, int NT, bool TWOX, bool UNIT, bool CG>
__device__ __forceinline__ double stencil_row_of_kind(const SpmvArgs &a, const StencilParams &p, int kind, int cself) {
    constexpr int B = NT == 1 ? 19 : 10;
    if (!P2) return stencil_row<7, 7, NT, TWOX, UNIT, CG>(a, p, 0, cself);
    switch (kind) {  // warp-uniform
    case 0: return stencil_row<19, B, NT, TWOX, UNIT, CG>(a, p, 0, cself);
    case 1: return stencil_row<9, 9, NT, TWOX, UNIT, CG>(a, p, 1, cself);
    case 2: return stencil_row<9, 9, NT, TWOX, UNIT, CG>(a, p, 2, cself);
    default: return stencil_row<9, 9, NT, TWOX, UNIT, CG>(a, p, 3, cself);
    }
}
__device__ __forceinline__ void prefetch_l2(const double *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// CGEPI: the CG iteration's launch (y = A d, d . y) gets a two-instruction epilogue instead of the generic one
template <bool P2, int NT, bool TWOX, bool UNIT, bool CGEPI>
__global__ void __launch_bounds__(kThreads, (P2 && !(UNIT && NT == 1)) ? 3 : 4) k_spmv_st(
    SpmvArgs a, const __grid_constant__ StencilParams p) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (a.skip_flag && *a.skip_flag != 0) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (kThreads / 32);
    const int warp0 = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    double dots[2] = {0.0, 0.0};
    bool halo_ready = a.halo_wait_seq == 0;
    // ---- phase 1: stencil slices, in the order of st_walk.  The host lays the walk out in tiles (the warps
    // of a block take slices of one column range in neighbouring DoF lines and kinds, so their x gathers
    // share L1 lines), rotates it for several ranks so that the ghost-reading slices at both ends of the
    // rank's rows are reached mid-kernel (their halo has long arrived by then), and terminates it with
    // sentinel entries, so a warp just follows a pointer.  Every x entry is some row's own column: each
    // warp asks L2 for the own-column segment of the slice it will take kStencilPrefetch trips later, so
    // the gathers find their lines in L2 instead of paying the HBM latency with a line or two in flight.
    {
        const int2 *w = a.A.st_walk + warp0;
        const size_t step = (size_t)nwarps;
        int2 wm = *w;                                  // (slice | ghost << 27 | kind << 28, first row or -1)
        int pf_row = w[step * kStencilPrefetch].y;
        while (wm.x >= 0) {
            w += step;
            const int2 nwm = *w;
            if (pf_row >= 0) prefetch_l2(a.t[0].xa + (pf_row + lane + a.A.own_off));
            pf_row = w[step * kStencilPrefetch].y;
            const int kind = wm.x >> 28;
            int r = wm.y + lane;
            if (wm.y < 0) r = a.A.row_of[(wm.x & 0x07ffffff) * kSlice + lane];
            const int cself = r + a.A.own_off;
            double s;
            if (wm.x & (1 << 27)) {  // reads ghost entries (several ranks): wait for the halo, gather through L2
                if (!halo_ready) {
                    halo_wait(a, lane);
                    halo_ready = true;
                }
                s = stencil_row_of_kind<P2, NT, TWOX, UNIT, true>(a, p, kind, cself);
            } else
                s = stencil_row_of_kind<P2, NT, TWOX, UNIT, false>(a, p, kind, cself);
            if (CGEPI) {
                a.y[r] = s;
                dots[0] += s * a.dotv[r];
            } else
                spmv_epilogue(a, r, s, dots);
            wm = nwm;
        }
    }
    // ---- phase 2: the slices kept in SELL form -------------------------------------------------------
    constexpr int CS = 4;
    for (int q = warp0; q < a.A.n_sell; q += nwarps) {
        const int slice = a.A.sell_list[q];
        const bool ghosty = a.pc.enabled && (slice < a.ghost_lo_slices || slice >= a.ghost_hi_slice0);
        if (ghosty && !halo_ready) {
            halo_wait(a, lane);
            halo_ready = true;
        }
        const uint32_t b = a.A.slice_ptr[slice], e = a.A.slice_ptr[slice + 1];
        const int r = a.A.row_of[slice * kSlice + lane];
        const int len = (int)((e - b) >> 5);
        const uint32_t base = b + lane;
        double s = 0.0;
        for (int k0 = 0; k0 < len; k0 += CS) {
            int c[CS];
            double v[NT][CS];
#pragma unroll
            for (int k = 0; k < CS; ++k) {
                const uint32_t qq = base + (uint32_t)min(k0 + k, len - 1) * kSlice;
                c[k] = ld_stream(&a.A.col[qq]);
#pragma unroll
                for (int t = 0; t < NT; ++t) v[t][k] = ld_stream(&a.t[t].val[qq]);
            }
#pragma unroll
            for (int k = 0; k < CS; ++k) {
                if (k0 + k < len) {  // warp-uniform
                    double prod = 0.0;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const double xa = ghosty ? __ldcg(&a.t[t].xa[c[k]]) : a.t[t].xa[c[k]];
                        double xv = __dmul_rn(a.t[t].ca, xa);
                        if (TWOX && a.t[t].xb)
                            xv = __dadd_rn(xv, __dmul_rn(a.t[t].cb, ghosty ? __ldcg(&a.t[t].xb[c[k]]) : a.t[t].xb[c[k]]));
                        const double term = __dmul_rn(a.t[t].coef, __dmul_rn(v[t][k], xv));
                        prod = t == 0 ? term : __dadd_rn(prod, term);
                    }
                    s = __dadd_rn(s, prod);
                }
            }
        }
        if (r >= 0) spmv_epilogue(a, r, s, dots);
    }
    spmv_tail(a, dots);
}

__global__ void k_zero_rows(int nb, const int32_t *brow, double *vec) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) vec[brow[b]] = 0.0;
}

// ---- K6: PCG with device-resident scalars (deal.II SolverCG) ----------------------------
// after the residual SpMV + all-reduce: iteration_status(0, res0)
__global__ void k_cg_start(CgScalars *S) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double res0 = sqrt(S->gg);
    S->res0 = res0;
    S->res = res0;
    S->reduced_tol = res0 * S->reduce;
    S->it = 0;
    S->gh[0] = S->gh_new;
    S->status[0] = (res0 <= S->reduced_tol || res0 <= S->tol) ? 1 : 0;
    S->status[1] = 0;
}
// Iteration k, parity = k & 1.  alpha = gh / dAd ; g += alpha h ; res^2 = g.g ; h = D^-1 g ; gh' = g.h
// (one pass: 3 reads, 2 writes).  The x update of this iteration (x += alpha d) is done by k_cg_direction,
// which reads d anyway.  Four independent elements per thread and trip keep ~100 bytes in flight per thread.
__global__ void __launch_bounds__(kThreads) k_cg_update(int n, int parity, CgScalars *S, double *__restrict__ g,
                                                        double *__restrict__ h, const double *__restrict__ dinv,
                                                        CgSumIo in, double *out_partials, unsigned *counter,
                                                        PeerComm pc, unsigned long long out_seq, int out_mode) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (S->status[parity] != 0) return;
    double dAd[1];
    consume_sums<1>(in, pc, &S->dAd, dAd);
    const double alpha = S->gh[parity] / dAd[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) S->alpha = alpha;
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * kThreads * 4;
    for (int i0 = blockIdx.x * kThreads * 4 + threadIdx.x; i0 < n; i0 += stride) {
        double gv[4], hv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kThreads;
            if (i < n) {
                gv[u] = g[i];
                hv[u] = h[i];
                dv[u] = dinv ? dinv[i] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kThreads;
            if (i < n) {
                const double gi = gv[u] + alpha * hv[u];
                g[i] = gi;
                acc[0] += gi * gi;
                if (dinv) {  // Jacobi fused here; with the multigrid preconditioner h is produced by the V-cycle
                    const double hi = dv[u] * gi;
                    h[i] = hi;
                    acc[1] += gi * hi;
                }
            }
        }
    }
    produce_sums<2>(acc, out_mode, out_partials, counter, &S->gg);
}
// x += alpha d ; iteration_status(k + 1, res) ; beta = gh'/gh ; d = beta d - h   (3 reads, 2 writes).
// Block 0 records the outcome in the other parity's slots: nothing it writes is read by this kernel.
__global__ void __launch_bounds__(kThreads) k_cg_direction(int n, int k, CgScalars *S, double *__restrict__ x,
                                                           double *__restrict__ d, const double *__restrict__ h,
                                                           CgSumIo in, int gh_scalar, unsigned *counter, PeerComm pc,
                                                           unsigned long long halo_seq) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    const int parity = k & 1;
    const int st = S->status[parity];
    if (st != 0) {  // finished earlier: hand the outcome on to the slot the next iteration reads
        if (blockIdx.x == 0 && threadIdx.x == 0) S->status[parity ^ 1] = st;
        return;
    }
    double tot[2];
    consume_sums<2>(in, pc, &S->gg, tot);
    const double gh_new = gh_scalar ? S->gh_new : tot[1];
    const double gh_old = S->gh[parity];
    const double alpha = S->alpha;  // the step length of this iteration (k_cg_update)
    const double res = sqrt(fabs(tot[0]));
    const int it = k + 1;
    int status = 0;
    if (res <= S->reduced_tol || res <= S->tol) status = 1;
    else if (it >= S->maxit || isnan(res)) status = 2;
    const double beta = gh_new / gh_old;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S->it = it;
        S->res = res;
        S->gh[parity ^ 1] = gh_new;
        S->status[parity ^ 1] = status;
    }
    // Several ranks: the first / last owned block is the neighbours' ghost block.  Those elements are updated
    // first and stored into the neighbours' vectors at once; every block then checks in (system-scope fence +
    // ticket) and the last one to arrive raises the halo flags -- a few microseconds into the kernel instead
    // of at its end, so the neighbours' next SpMV never waits and this kernel has no tail.
    int mid0 = 0, mid1 = n;
    if (pc.enabled && status == 0) {
        const int hi0 = n - pc.hi_count;
        mid0 = pc.lo_count;
        mid1 = hi0 > mid0 ? hi0 : mid0;
        for (int part = 0; part < 2; ++part) {
            const int b = part == 0 ? 0 : mid1, e = part == 0 ? mid0 : n;
            for (int i = b + blockIdx.x * kThreads + threadIdx.x; i < e; i += gridDim.x * kThreads) {
                const double dold = d[i];
                x[i] += alpha * dold;
                const double di = beta * dold - h[i];
                d[i] = di;
                if (pc.d_lo && i < pc.lo_count) pc.d_lo[i] = di;
                if (pc.d_hi && i >= hi0) pc.d_hi[i - hi0] = di;
            }
        }
        if (last_block_done(counter, true) && threadIdx.x == 0) {
            __threadfence_system();
            if (pc.rank > 0) st_release_sys(&pc.box[pc.rank - 1]->halo_flag[1], halo_seq);
            if (pc.rank < pc.nranks - 1) st_release_sys(&pc.box[pc.rank + 1]->halo_flag[0], halo_seq);
        }
    }
    const int stride = gridDim.x * kThreads * 4;
    for (int i0 = mid0 + blockIdx.x * kThreads * 4 + threadIdx.x; i0 < mid1; i0 += stride) {
        double xv[4], dv[4], hv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kThreads;
            if (i < mid1) {
                xv[u] = x[i];
                dv[u] = d[i];
                hv[u] = status == 0 ? h[i] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kThreads;
            if (i < mid1) {
                x[i] = xv[u] + alpha * dv[u];
                if (status == 0) d[i] = beta * dv[u] - hv[u];
            }
        }
    }
}

// ---- multigrid V-cycle pieces -----------------------------------------------------------------------
// Transfers are FE interpolation between nested spaces, written as closed-form stencils: a fine P1
// vertex is a coarse vertex (weight 1) or the midpoint of a coarse edge -- horizontal, vertical or the
// diagonal (i+1,j)-(i,j+1) of the structured triangulation (weights 1/2, 1/2); a P2 edge DoF is the
// midpoint of its edge.  Restriction is the transpose, with Dirichlet rows of the coarse level zeroed.
__global__ void k_scale_rows(int n, double omega, const double *__restrict__ dinv, const double *__restrict__ b,
                             double *__restrict__ x, const int *skip_flag) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (skip_flag && *skip_flag != 0) return;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = omega * (dinv[i] * b[i]);
}
// All four transfers are written for a rank's strip: one thread per owned DoF of the target level; the
// source vector is in local layout (ghost blocks valid: the caller exchanged its halo).  Restriction needs
// the fine level's upper ghost block, prolongation the coarse level's lower one (coarse quad row J covers
// the fine quad rows 2J, 2J+1, so both levels split at the same physical lines).
__global__ void k_prolong_add_p1(Layout Lf, Layout Lc, const double *__restrict__ ec, double *__restrict__ xf,
                                 const int *skip_flag) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (skip_flag && *skip_flag != 0) return;
    const Mesh &mf = Lf.mesh, &mc = Lc.mesh;
    const SlotRange s = owned_slots(Lf);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(Lf, s)) return;
    int I, J, kind;
    decode_slot(Lf, s, t, I, J, kind);
    const int64_t dof = dof_V(mf, I, J);
    if (dof < Lf.row0 || dof >= Lf.row0 + Lf.nown) return;
    auto c = [&](int i, int j) { return ec[dof_V(mc, i, j) - Lc.col0]; };
    double v;
    if (!(I & 1) && !(J & 1)) v = c(I / 2, J / 2);
    else if ((I & 1) && !(J & 1)) v = 0.5 * (c((I - 1) / 2, J / 2) + c((I + 1) / 2, J / 2));
    else if (!(I & 1)) v = 0.5 * (c(I / 2, (J - 1) / 2) + c(I / 2, (J + 1) / 2));
    else v = 0.5 * (c((I + 1) / 2, (J - 1) / 2) + c((I - 1) / 2, (J + 1) / 2));
    xf[dof - Lf.col0] += v;
}
__global__ void k_restrict_p1(Layout Lf, Layout Lc, const double *__restrict__ rf, double *__restrict__ bc,
                              const int *skip_flag) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (skip_flag && *skip_flag != 0) return;
    const Mesh &mf = Lf.mesh, &mc = Lc.mesh;
    const SlotRange s = owned_slots(Lc);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(Lc, s)) return;
    int i, j, kind;
    decode_slot(Lc, s, t, i, j, kind);
    const int64_t dof = dof_V(mc, i, j);
    if (dof < Lc.row0 || dof >= Lc.row0 + Lc.nown) return;
    auto f = [&](int I, int J) { return rf[dof_V(mf, I, J) - Lf.col0]; };
    double sum = 0.0;
    if (i > 0 && i < mc.nx && j > 0 && j < mc.ny) {
        const int I = 2 * i, J = 2 * j;
        sum = f(I, J) + 0.5 * (f(I - 1, J) + f(I + 1, J) + f(I, J - 1) + f(I, J + 1) + f(I + 1, J - 1) + f(I - 1, J + 1));
    }
    bc[dof - Lc.row0] = sum;
}
__global__ void k_prolong_add_p2p1(Layout Lf, Layout Lc, const double *__restrict__ ec, double *__restrict__ xf,
                                   const int *skip_flag) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (skip_flag && *skip_flag != 0) return;
    const Mesh &mf = Lf.mesh, &mc = Lc.mesh;
    const SlotRange s = owned_slots(Lf);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(Lf, s)) return;
    int i, j, kind;
    decode_slot(Lf, s, t, i, j, kind);
    const int64_t dof = entity_dof_internal(mf, i, j, kind);
    if (dof < Lf.row0 || dof >= Lf.row0 + Lf.nown) return;
    auto c = [&](int ci, int cj) { return ec[dof_V(mc, ci, cj) - Lc.col0]; };
    double v;
    if (kind == 0) v = c(i, j);
    else if (kind == 1) v = 0.5 * (c(i, j) + c(i + 1, j));
    else if (kind == 2) v = 0.5 * (c(i, j) + c(i, j + 1));
    else v = 0.5 * (c(i + 1, j) + c(i, j + 1));
    xf[dof - Lf.col0] += v;
}
__global__ void k_restrict_p2p1(Layout Lf, Layout Lc, const double *__restrict__ rf, double *__restrict__ bc,
                                const int *skip_flag) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (skip_flag && *skip_flag != 0) return;
    const Mesh &mf = Lf.mesh, &mc = Lc.mesh;
    const SlotRange s = owned_slots(Lc);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slot_count(Lc, s)) return;
    int i, j, kind;
    decode_slot(Lc, s, t, i, j, kind);
    const int64_t dof = dof_V(mc, i, j);
    if (dof < Lc.row0 || dof >= Lc.row0 + Lc.nown) return;
    double sum = 0.0;
    if (i > 0 && i < mc.nx && j > 0 && j < mc.ny) {
        const int64_t o = Lf.col0;
        sum = rf[idof_V(mf, i, j) - o] +
              0.5 * (rf[idof_B(mf, i - 1, j) - o] + rf[idof_B(mf, i, j) - o] + rf[idof_L(mf, i, j - 1) - o] +
                     rf[idof_L(mf, i, j) - o] + rf[idof_D(mf, i - 1, j) - o] + rf[idof_D(mf, i, j - 1) - o]);
    }
    bc[dof - Lc.row0] = sum;
}
__global__ void __launch_bounds__(kThreads) k_dot_gz(int n, const double *__restrict__ g, const double *__restrict__ z,
                                                     double *d, double *partials, unsigned *counter, double *result,
                                                     const int *skip_flag) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (skip_flag && *skip_flag != 0) return;
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double zi = z[i];
        acc[0] += g[i] * zi;
        if (d) d[i] = -zi;
    }
    if (grid_sum<1>(acc, partials, counter) && threadIdx.x == 0) result[0] = acc[0];
}

// ---- K7: fused Newmark vector updates (src/WaveNewmark.cpp:121-126, :264-278, :429-430) --------------
// predictor, in place: u <- z = u + dt v + dt^2(1/2-beta) a ;  v <- v + dt(1-gamma) a
__global__ void __launch_bounds__(kThreads) k_newmark_predict(int n, double dt, double c1, double c2,
                                                              double *__restrict__ u, double *__restrict__ v,
                                                              const double *__restrict__ a) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ui = u[i], vi = v[i], ai = a[i];
        u[i] = (ui + dt * vi) + c1 * ai;
        v[i] = vi + c2 * ai;
    }
}
// corrector: u <- z + beta dt^2 a+ ; v <- v + dt gamma a+ ; ||u||^2, ||v||^2
__global__ void __launch_bounds__(kThreads) k_newmark_correct(int n, double cu, double cv,
                                                              double *__restrict__ u, double *__restrict__ v,
                                                              const double *__restrict__ a, double *partials,
                                                              unsigned *counter, double *result) {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ai = a[i];
        const double ui = u[i] + cu * ai, vi = v[i] + cv * ai;
        u[i] = ui;
        v[i] = vi;
        acc[0] += ui * ui;
        acc[1] += vi * vi;
    }
    if (grid_sum<2>(acc, partials, counter) && threadIdx.x == 0) { result[0] = acc[0]; result[1] = acc[1]; }
}
__global__ void __launch_bounds__(kThreads) k_norms2(int n, const double *__restrict__ u,
                                                     const double *__restrict__ v, double *partials,
                                                     unsigned *counter, double *result) {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        acc[0] += u[i] * u[i];
        acc[1] += v[i] * v[i];
    }
    if (grid_sum<2>(acc, partials, counter) && threadIdx.x == 0) { result[0] = acc[0]; result[1] = acc[1]; }
}
__global__ void __launch_bounds__(kThreads) k_copy(int n, const double *__restrict__ src, double *__restrict__ dst) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
__global__ void __launch_bounds__(kThreads) k_fill(int64_t n, double value, double *__restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

// ---- error norms: compute_error / compute_relative_error (src/WaveEquationBase.cpp:367-423) ---------
// [deal.II] integrate_difference with QGaussSimplex(r+2); per-cell numerator rounded to float (:384),
// exact-solution gradient by centred differences h = 1e-8 (FunctionParser is an AutoDerivativeFunction).
template <int R>
__global__ void __launch_bounds__(128) k_errors(Layout L, const Program *sol, Quadrature Q, double t,
                                                const double *u, double *partials, unsigned *counter,
                                                double *result) {
    constexpr int DPC = R == 1 ? 3 : 6;
    const int64_t c_begin = 2LL * L.jq0 * L.mesh.nx, c_end = 2LL * L.jq1 * L.mesh.nx;
    const int64_t cell = c_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (cell < c_end) {
        double X0, Y0, sx, sy;
        cell_geometry(L.mesh, cell, X0, Y0, sx, sy);
        const double det = sx * sy, adet = fabs(det);
        const double ix = sy / det, iy = sx / det;
        int64_t d[6];
        cell_dofs_internal(L.mesh, cell, d);
        double ul[DPC];
#pragma unroll
        for (int a = 0; a < DPC; ++a) ul[a] = u[d[a] - L.col0];
        const double hfd = 1e-8;
        double dl2 = 0, dsemi = 0, xl2 = 0, xsemi = 0;
        for (int q = 0; q < Q.nq; ++q) {
            double phi[DPC], dxi[DPC], deta[DPC];
            shape_values(R, Q.xi[q], Q.eta[q], phi);
            shape_grads(R, Q.xi[q], Q.eta[q], dxi, deta);
            const double JxW = Q.w[q] * adet;
            const double xq = X0 + sx * Q.xi[q], yq = Y0 + sy * Q.eta[q];
            double uh = 0, ux = 0, uy = 0;
#pragma unroll
            for (int a = 0; a < DPC; ++a) {
                uh += ul[a] * phi[a];
                ux += ul[a] * (ix * dxi[a]);
                uy += ul[a] * (iy * deta[a]);
            }
            const double ue = eval(sol, xq, yq, t);
            const double uex = (eval(sol, xq + hfd, yq, t) - eval(sol, xq - hfd, yq, t)) / (2 * hfd);
            const double uey = (eval(sol, xq, yq + hfd, t) - eval(sol, xq, yq - hfd, t)) / (2 * hfd);
            dl2 += (uh - ue) * (uh - ue) * JxW;
            dsemi += ((ux - uex) * (ux - uex) + (uy - uey) * (uy - uey)) * JxW;
            xl2 += ue * ue * JxW;
            xsemi += (uex * uex + uey * uey) * JxW;
        }
        const float fl2 = (float)sqrt(dl2), fh1 = (float)sqrt(dl2 + dsemi);
        acc[0] = (double)fl2 * (double)fl2;
        acc[1] = (double)fh1 * (double)fh1;
        acc[2] = xl2;
        acc[3] = xl2 + xsemi;
    }
    if (grid_sum<4>(acc, partials, counter) && threadIdx.x == 0) {
        result[0] = acc[0]; result[1] = acc[1]; result[2] = acc[2]; result[3] = acc[3];
    }
}

// log_point_probe: u_h at a point (src/WaveEquationBase.cpp:170-206); the rank owning the quad row
// evaluates, the others contribute 0 (the reference's MPI_Reduce SUM).
__global__ void k_probe(Layout L, double px, double py, const double *u, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Mesh &m = L.mesh;
    double val = 0.0;
    int i = (int)floor((px - m.x0) / m.dx), j = (int)floor((py - m.y0) / m.dy);
    if (i >= m.nx) i = m.nx - 1;
    if (j >= m.ny) j = m.ny - 1;
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    // the oracle walks cells in order and takes the first that contains the point: prefer the
    // lower-index candidates (left / lower quad, T0 before T1) when the point sits on an edge
    const double eps = 1e-12;
    bool found = false;
    for (int jj = (j > 0 ? j - 1 : 0); jj <= j && !found; ++jj)
        for (int ii = (i > 0 ? i - 1 : 0); ii <= i && !found; ++ii)
            for (int tt = 0; tt < 2 && !found; ++tt) {
                const int64_t cell = 2 * ((int64_t)jj * m.nx + ii) + tt;
                double X0, Y0, sx, sy;
                cell_geometry(m, cell, X0, Y0, sx, sy);
                const double xi = (px - X0) / sx, eta = (py - Y0) / sy;
                if (xi >= -eps && eta >= -eps && xi + eta <= 1.0 + eps) {
                    found = true;
                    if (jj >= L.jq0 && jj < L.jq1) {
                        double phi[6];
                        int64_t d[6];
                        shape_values(m.r, xi, eta, phi);
                        cell_dofs_internal(m, cell, d);
                        for (int a = 0; a < dofs_per_cell(m.r); ++a) val += u[d[a] - L.col0] * phi[a];
                    }
                }
            }
    *out = val;
}

__global__ void __launch_bounds__(kThreads) k_flush(double *buf, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = buf[i] * 0.5 + 1.0;
}

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }
// grid-stride streams: a multiple of the SM count, capped by the work available
inline int stream_blocks(int64_t n) {
    const int64_t need = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kSMs * 8;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace

int reduction_blocks(int n) { return stream_blocks(n); }

static inline void note_launch(const Launcher &l) {
    if (l.count) ++*l.count;
    if (l.error && *l.error == cudaSuccess) *l.error = cudaPeekAtLastError();
}
#define WV_LAUNCH(l, kernel, grid, block, smem, ...)                  \
    do {                                                              \
        kernel<<<(grid), (block), (smem), (l).stream>>>(__VA_ARGS__); \
        note_launch(l);                                               \
    } while (0)

// Launch through cudaLaunchKernelEx with the programmatic-stream-serialization attribute as an option
// (WAVE_PDL=1): the kernels start with cudaGridDependencySynchronize(), so their blocks may be scheduled
// while the previous kernel of the stream drains.  Round 1 measured -11 % step time with it at c2 size; with
// the round-2 kernels (no last-block tails, consumer-side sums) it no longer pays: measured on B200 it costs
// 1-2 % at 2-8 M rows, 9 % at 67 M rows (early-resident blocks of the next kernel take registers and L1 from
// the running one) and 16 us per CG iteration with several ranks, so it is off by default.
template <class... KArgs, class... Args>
static void launch_pdl(const Launcher &l, void (*kernel)(KArgs...), int grid, int block, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = l.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    static const int pdl = (std::getenv("WAVE_PDL") && std::atoi(std::getenv("WAVE_PDL")) != 0) ? 1 : 0;
    attr[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (l.count) ++*l.count;
    if (l.error && *l.error == cudaSuccess) *l.error = e != cudaSuccess ? e : cudaPeekAtLastError();
}

void launch_row_lengths(const Launcher &l, const Layout &L, uint32_t *rowlen) {
    const int64_t n = slot_count(L, owned_slots(L));
    WV_LAUNCH(l, k_row_lengths, blocks_for(n, 128), 128, 0, L, rowlen);
}
void launch_window_sort(const Launcher &l, int nown, int nslots, const uint32_t *rowlen, int32_t *row_of,
                        int32_t *slot_of) {
    WV_LAUNCH(l, k_window_sort, nslots / kWindow, kWindow, 0, nown, rowlen, row_of, slot_of);
}
void launch_slice_sizes(const Launcher &l, int nslices, const int32_t *row_of, const uint32_t *rowlen,
                        uint32_t *slice_cnt) {
    WV_LAUNCH(l, k_slice_sizes, blocks_for((int64_t)nslices * 32, kThreads), kThreads, 0, nslices, row_of, rowlen,
              slice_cnt);
}
void launch_fill_int(const Launcher &l, int64_t n, int32_t value, int32_t *dst) {
    if (n <= 0) return;
    WV_LAUNCH(l, k_fill_int, stream_blocks(n), kThreads, 0, n, value, dst);
}
void launch_fill_cols(const Launcher &l, const Layout &L, const Sell &A, int32_t *col) {
    const int64_t n = slot_count(L, owned_slots(L));
    WV_LAUNCH(l, k_fill_cols, blocks_for(n, 128), 128, 0, L, A, col);
}
void launch_canonical_lengths(const Launcher &l, const Layout &L, const Sell &A, const int32_t *c2i, uint32_t *len) {
    WV_LAUNCH(l, k_canonical_lengths, blocks_for(L.nown, kThreads), kThreads, 0, L, A, c2i, len);
}
void launch_export_csr(const Launcher &l, const Layout &L, const Sell &A, const int32_t *c2i, const int32_t *i2c,
                       const uint32_t *rowptr_c, const double *val, double *csr_val, int32_t *csr_col) {
    WV_LAUNCH(l, k_export_csr, blocks_for(L.nown, kThreads), kThreads, 0, L, A, c2i, i2c, rowptr_c, val, csr_val,
              csr_col);
}
void launch_build_perm(const Launcher &l, const Layout &L, int32_t *c2i, int32_t *i2c) {
    const int64_t n = slot_count(L, local_slots(L));
    WV_LAUNCH(l, k_build_perm, blocks_for(n, 128), 128, 0, L, c2i, i2c);
}
void launch_gather(const Launcher &l, int n, const int32_t *map, const double *src, double *dst) {
    if (n <= 0) return;
    WV_LAUNCH(l, k_gather, stream_blocks(n), kThreads, 0, n, map, src, dst);
}
void launch_assemble(const Launcher &l, const Layout &L, const Program *c, const Quadrature *q, const Sell &A,
                     double *M, double *K) {
    const int64_t n = slot_count(L, owned_slots(L));
    if (L.mesh.r == 1) WV_LAUNCH(l, k_assemble<1>, blocks_for(n, 128), 128, 0, L, c, *q, A, M, K);
    else WV_LAUNCH(l, k_assemble<2>, blocks_for(n, 128), 128, 0, L, c, *q, A, M, K);
}
void launch_stencil_tables(const Launcher &l, const Layout &L, const Program *c, const Quadrature *q, int32_t *st_meta,
                           double *tabM, double *tabK) {
    if (L.mesh.r == 1) WV_LAUNCH(l, k_stencil_tables<1>, 1, 32, 0, L, c, *q, st_meta, tabM, tabK);
    else WV_LAUNCH(l, k_stencil_tables<2>, 1, 32, 0, L, c, *q, st_meta, tabM, tabK);
}
void launch_stencil_classify(const Launcher &l, const Layout &L, const Sell &A, const double *M, const double *K,
                             const int32_t *st_meta, const double *tabM, const double *tabK, int8_t *row_kind,
                             int2 *slice_info, unsigned long long *counts) {
    const int64_t n = slot_count(L, owned_slots(L));
    WV_LAUNCH(l, k_classify_rows, blocks_for(n, 128), 128, 0, L, A, M, K, st_meta, tabM, tabK, row_kind);
    WV_LAUNCH(l, k_classify_slices, blocks_for((int64_t)A.nslices * 32, kThreads), kThreads, 0, A, row_kind, slice_info,
              counts);
}
void launch_axpy_vals(const Launcher &l, int64_t nnz, const double *M, const double *K, double s, double *out) {
    WV_LAUNCH(l, k_axpy_vals, stream_blocks(nnz), kThreads, 0, nnz, M, K, s, out);
}
void launch_find_d0(const Launcher &l, const Layout &L, const Sell &A, const double *val, double *d0) {
    WV_LAUNCH(l, k_find_d0, 1, 32, 0, L, A, val, d0);
}
void launch_bc_rows(const Launcher &l, const Layout &L, int nb, const int32_t *brow, const Sell &A, double *val,
                    const double *d0) {
    if (nb <= 0) return;
    WV_LAUNCH(l, k_bc_rows, blocks_for(nb, 128), 128, 0, L, nb, brow, A, val, d0);
}
void launch_dinv(const Launcher &l, const Layout &L, const Sell &A, const double *val, int identity, double *dinv) {
    WV_LAUNCH(l, k_dinv, blocks_for(L.nown, kThreads), kThreads, 0, L, A, val, identity, dinv);
}
void launch_interpolate(const Launcher &l, const Layout &L, const Program *p, double t, double *vec,
                        double *sx, double *sy) {
    const int64_t n = slot_count(L, local_slots(L));
    WV_LAUNCH(l, k_interpolate, blocks_for(n, 128), 128, 0, L, p, t, vec, sx, sy);
}
// cells whose load vectors the owned rows need: the owned quad rows and the quad row above them
int64_t forcing_cells(const Layout &L) {
    const int jtop = L.jq1 < L.mesh.ny ? L.jq1 : L.mesh.ny - 1;
    return 2LL * (jtop + 1 - L.jq0) * L.mesh.nx;
}
void launch_forcing(const Launcher &l, const Layout &L, const Program *f, const Quadrature *q, double t_np1,
                    double t_n, double w_np1, double w_n, int two_levels, double *cellvec, double *fvec) {
    const int64_t nc = forcing_cells(L);
    const int64_t nr = slot_count(L, owned_slots(L));
    if (L.mesh.r == 1) {
        WV_LAUNCH(l, k_forcing_cells<1>, blocks_for(nc, 128), 128, 0, L, f, *q, t_np1, t_n, w_np1, w_n, two_levels,
                  cellvec);
        WV_LAUNCH(l, k_forcing_gather<1>, blocks_for(nr, 128), 128, 0, L, cellvec, fvec);
    } else {
        WV_LAUNCH(l, k_forcing_cells<2>, blocks_for(nc, 128), 128, 0, L, f, *q, t_np1, t_n, w_np1, w_n, two_levels,
                  cellvec);
        WV_LAUNCH(l, k_forcing_gather<2>, blocks_for(nr, 128), 128, 0, L, cellvec, fvec);
    }
}
void launch_bc_values(const Launcher &l, int mode, int nb, const int32_t *brow, const double *bx,
                      const double *by, const Program *g, double t, double dt, double beta_dt2,
                      const double *z_own, double *x_own, double *rhs, const double *d0) {
    if (nb <= 0) return;
    WV_LAUNCH(l, k_bc_values, blocks_for(nb, 128), 128, 0, mode, nb, brow, bx, by, g, t, dt, beta_dt2, z_own,
              x_own, rhs, d0);
}
template <class Kernel>
static int persistent_grid(Kernel kernel, int64_t blocks_needed) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0) != cudaSuccess || occ < 1) occ = 4;
    const int64_t cap = (int64_t)kSMs * occ;
    return (int)(blocks_needed < cap ? (blocks_needed > 0 ? blocks_needed : 1) : cap);
}
template <int NT, bool TWOX, int CH, int MINB>
static int spmv_blocks_t(const SpmvArgs &a) {
    static int grid_cap = 0;
    if (!grid_cap) grid_cap = persistent_grid(k_spmv<NT, TWOX, CH, MINB>, 1 << 30);
    const int64_t need = blocks_for((int64_t)a.A.nslices * kSlice, kThreads);
    return (int)std::min<int64_t>(need, grid_cap);
}
template <int NT, bool TWOX, int CH, int MINB>
static void launch_spmv_t(const Launcher &l, const SpmvArgs &a) {
    launch_pdl(l, k_spmv<NT, TWOX, CH, MINB>, spmv_blocks_t<NT, TWOX, CH, MINB>(a), kThreads, a);
}

template <bool P2, int NT, bool TWOX, bool UNIT, bool CGEPI = false>
static int spmv_st_blocks_t(const SpmvArgs &a) {
    static int grid_cap = 0;
    if (!grid_cap) grid_cap = persistent_grid(k_spmv_st<P2, NT, TWOX, UNIT, CGEPI>, 1 << 30);
    const int64_t need = blocks_for((int64_t)a.A.nslices * kSlice, kThreads);
    return (int)std::min<int64_t>(need, grid_cap);
}
template <bool P2, int NT, bool TWOX, bool UNIT, bool CGEPI = false>
static void launch_spmv_st_t(const Launcher &l, const SpmvArgs &a) {
    StencilParams p;  // host tables -> kernel parameter (constant bank)
    std::memcpy(p.off, a.A.st_meta, sizeof p.off);
    for (int t = 0; t < 2; ++t)
        if (t < NT) std::memcpy(p.val[t], a.t[t].tab, sizeof p.val[t]);
        else std::memset(p.val[t], 0, sizeof p.val[t]);
    launch_pdl(l, k_spmv_st<P2, NT, TWOX, UNIT, CGEPI>, spmv_st_blocks_t<P2, NT, TWOX, UNIT, CGEPI>(a), kThreads, a, p);
}
static bool use_stencil(const SpmvArgs &a) {
    if (!a.A.slice_info || !a.t[0].tab) return false;
    return a.t[1].val == nullptr || a.t[1].tab != nullptr;
}
static bool unit_coefficients(const SpmvArgs &a) {
    return a.t[1].val == nullptr && a.t[0].xb == nullptr && a.t[0].ca == 1.0 && a.t[0].coef == 1.0;
}
// coefficient -1 on a single vector (-K z of the right-hand side, b - S x of the multigrid smoother): every
// term -(v x) is the exact negation of (v x) and rounding is symmetric, so the row sum is the exact negation
// of the unit-coefficient sum -- the stencil kernel runs its unit variant and the epilogue flips the sign
static bool unit_negative(const SpmvArgs &a) {
    return a.t[1].val == nullptr && a.t[0].xb == nullptr && a.t[0].ca == 1.0 && a.t[0].coef == -1.0;
}
// the CG iteration's launch: y = A d and d . y, nothing else in the epilogue
static bool cg_epilogue(const SpmvArgs &a) {
    return unit_coefficients(a) && a.y && a.dot_mode == 1 && !a.add0 && !a.add1 && !a.jac_x && !a.h_out;
}
// the grid of the variant the CG iteration launches (single term, single vector, unit coefficients)
int spmv_launch_blocks(const SpmvArgs &a) {
    if (use_stencil(a))
        return a.A.chunk <= 7 ? spmv_st_blocks_t<false, 1, false, true, true>(a) : spmv_st_blocks_t<true, 1, false, true, true>(a);
    return a.A.chunk <= 7 ? spmv_blocks_t<1, false, 7, 4>(a) : spmv_blocks_t<1, false, 10, 2>(a);
}
// chunk = the dominant row length of the element: P1 rows hold 7 entries, P2 rows 19 / 9
void launch_spmv(const Launcher &l, const SpmvArgs &a) {
    const bool two_terms = a.t[1].val != nullptr;
    const bool twox = a.t[0].xb != nullptr || (two_terms && a.t[1].xb != nullptr);
    const bool p1 = a.A.chunk <= 7;
    if (use_stencil(a) && unit_negative(a)) {
        SpmvArgs b = a;
        b.t[0].coef = 1.0;
        b.negate = 1;
        if (p1) launch_spmv_st_t<false, 1, false, true>(l, b); else launch_spmv_st_t<true, 1, false, true>(l, b);
        return;
    }
    if (use_stencil(a)) {
        if (cg_epilogue(a)) { if (p1) launch_spmv_st_t<false, 1, false, true, true>(l, a); else launch_spmv_st_t<true, 1, false, true, true>(l, a); }
        else if (unit_coefficients(a)) { if (p1) launch_spmv_st_t<false, 1, false, true>(l, a); else launch_spmv_st_t<true, 1, false, true>(l, a); }
        else if (!two_terms && !twox) { if (p1) launch_spmv_st_t<false, 1, false, false>(l, a); else launch_spmv_st_t<true, 1, false, false>(l, a); }
        else if (!two_terms) { if (p1) launch_spmv_st_t<false, 1, true, false>(l, a); else launch_spmv_st_t<true, 1, true, false>(l, a); }
        else { if (p1) launch_spmv_st_t<false, 2, true, false>(l, a); else launch_spmv_st_t<true, 2, true, false>(l, a); }
        return;
    }
    // (CH, blocks/SM) measured on B200: P1 (7, 4) -> 1.05 of the measured copy bandwidth at Nel=4096,
    // P2 (10, 2) -> 0.97; higher occupancy with fewer loads in flight per lane was slower for both
    if (!two_terms && !twox) { if (p1) launch_spmv_t<1, false, 7, 4>(l, a); else launch_spmv_t<1, false, 10, 2>(l, a); }
    else if (!two_terms) { if (p1) launch_spmv_t<1, true, 7, 4>(l, a); else launch_spmv_t<1, true, 10, 2>(l, a); }
    else { if (p1) launch_spmv_t<2, true, 7, 4>(l, a); else launch_spmv_t<2, true, 10, 2>(l, a); }
}
void launch_zero_rows(const Launcher &l, int nb, const int32_t *brow, double *vec) {
    if (nb <= 0) return;
    WV_LAUNCH(l, k_zero_rows, blocks_for(nb, 128), 128, 0, nb, brow, vec);
}
void launch_cg_start(const Launcher &l, CgScalars *S) { WV_LAUNCH(l, k_cg_start, 1, 32, 0, S); }
void launch_scale_rows(const Launcher &l, int n, double omega, const double *dinv, const double *b, double *x,
                       const int *skip_flag) {
    launch_pdl(l, k_scale_rows, stream_blocks(n), kThreads, n, omega, dinv, b, x, skip_flag);
}
void launch_prolong_add_p1(const Launcher &l, const Layout &Lf, const Layout &Lc, const double *ec, double *xf,
                           const int *skip_flag) {
    const int64_t n = slot_count(Lf, owned_slots(Lf));
    launch_pdl(l, k_prolong_add_p1, blocks_for(n, kThreads), kThreads, Lf, Lc, ec, xf, skip_flag);
}
void launch_restrict_p1(const Launcher &l, const Layout &Lf, const Layout &Lc, const double *rf, double *bc,
                        const int *skip_flag) {
    const int64_t n = slot_count(Lc, owned_slots(Lc));
    launch_pdl(l, k_restrict_p1, blocks_for(n, kThreads), kThreads, Lf, Lc, rf, bc, skip_flag);
}
void launch_prolong_add_p2p1(const Launcher &l, const Layout &Lf, const Layout &Lc, const double *ec, double *xf,
                             const int *skip_flag) {
    const int64_t n = slot_count(Lf, owned_slots(Lf));
    launch_pdl(l, k_prolong_add_p2p1, blocks_for(n, kThreads), kThreads, Lf, Lc, ec, xf, skip_flag);
}
void launch_restrict_p2p1(const Launcher &l, const Layout &Lf, const Layout &Lc, const double *rf, double *bc,
                          const int *skip_flag) {
    const int64_t n = slot_count(Lc, owned_slots(Lc));
    launch_pdl(l, k_restrict_p2p1, blocks_for(n, kThreads), kThreads, Lf, Lc, rf, bc, skip_flag);
}
void launch_dot_gz(const Launcher &l, int n, const double *g, const double *z, double *d_or_null, double *partials,
                   unsigned *counter, double *result, const int *skip_flag) {
    launch_pdl(l, k_dot_gz, stream_blocks(n), kThreads, n, g, z, d_or_null, partials, counter, result, skip_flag);
}
// grid of the CG vector kernels: at most two blocks per SM (four elements per thread and trip), so that a
// 1 M-row vector still gives every thread ~14 elements and the consumer sums at most 296 partials
int cg_vector_blocks(int n) {
    const int64_t need = ((int64_t)n + kThreads * 4 - 1) / (kThreads * 4);
    const int64_t cap = (int64_t)kSMs * 2;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}
void launch_cg_update(const Launcher &l, int n, int parity, CgScalars *S, double *g, double *h, const double *dinv,
                      const CgSumIo &in, double *out_partials, unsigned *counter, const PeerComm &pc,
                      unsigned long long out_seq, int out_mode) {
    launch_pdl(l, k_cg_update, cg_vector_blocks(n), kThreads, n, parity, S, g, h, dinv, in, out_partials, counter, pc,
               out_seq, out_mode);
}
void launch_cg_direction(const Launcher &l, int n, int k, CgScalars *S, double *x, double *d, const double *h,
                         const CgSumIo &in, int gh_scalar, unsigned *counter, const PeerComm &pc,
                         unsigned long long halo_seq) {
    launch_pdl(l, k_cg_direction, cg_vector_blocks(n), kThreads, n, k, S, x, d, h, in, gh_scalar, counter, pc,
               halo_seq);
}
void launch_newmark_predict(const Launcher &l, int n, double dt, double c1, double c2, double *u, double *v,
                            const double *a) {
    WV_LAUNCH(l, k_newmark_predict, stream_blocks(n), kThreads, 0, n, dt, c1, c2, u, v, a);
}
void launch_newmark_correct(const Launcher &l, int n, double cu, double cv, double *u, double *v, const double *a,
                            double *partials, unsigned *counter, double *result) {
    WV_LAUNCH(l, k_newmark_correct, stream_blocks(n), kThreads, 0, n, cu, cv, u, v, a, partials, counter, result);
}
void launch_norms2(const Launcher &l, int n, const double *u, const double *v, double *partials,
                   unsigned *counter, double *result) {
    WV_LAUNCH(l, k_norms2, stream_blocks(n), kThreads, 0, n, u, v, partials, counter, result);
}
void launch_copy(const Launcher &l, int n, const double *src, double *dst) {
    WV_LAUNCH(l, k_copy, stream_blocks(n), kThreads, 0, n, src, dst);
}
void launch_fill(const Launcher &l, int64_t n, double value, double *dst) {
    if (n <= 0) return;
    WV_LAUNCH(l, k_fill, stream_blocks(n), kThreads, 0, n, value, dst);
}
void launch_errors(const Launcher &l, const Layout &L, const Program *sol, const Quadrature *q, double t,
                   const double *u_local, double *partials, unsigned *counter, double *result) {
    const int64_t n = 2LL * (L.jq1 - L.jq0) * L.mesh.nx;
    if (L.mesh.r == 1)
        WV_LAUNCH(l, k_errors<1>, blocks_for(n, 128), 128, 0, L, sol, *q, t, u_local, partials, counter, result);
    else
        WV_LAUNCH(l, k_errors<2>, blocks_for(n, 128), 128, 0, L, sol, *q, t, u_local, partials, counter, result);
}
void launch_probe(const Launcher &l, const Layout &L, double px, double py, const double *u_local, double *out) {
    WV_LAUNCH(l, k_probe, 1, 32, 0, L, px, py, u_local, out);
}
void launch_flush_l2(const Launcher &l, double *buf, int64_t n) {
    WV_LAUNCH(l, k_flush, stream_blocks(n), kThreads, 0, buf, n);
}

}  // namespace wv
namespace wv {
int spmv_grid_blocks(int nslices) { return (nslices + 7) / 8; }
}
