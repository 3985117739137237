// cg_fused.cu -- see cg_fused.cuh.
#include "cg_fused.cuh"

#include <cooperative_groups.h>

#include <climits>

namespace cg = cooperative_groups;

namespace wv {

namespace {

constexpr unsigned kFull = 0xffffffffu;
static_assert(kFusedThreads == kWindow, "thread t of a block holds slot t of each of its windows");

// Totals of NV per-thread values, identical in every thread of every block: block tree, one partial
// per block, grid barrier, then every block adds the partials in the same fixed order.
// `buf` holds gridDim.x * NV doubles and must not be the buffer of the previous call.
template <int NV>
__device__ __forceinline__ void grid_allsum(cg::grid_group &grid, double (&v)[NV], double *buf, double *s_red,
                                            double *s_bc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(kFull, v[k], off);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) s_red[warp * NV + k] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = s_red[lane * NV + k];  // kFusedThreads / 32 == 32 warps
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(kFull, x, off);
            if (lane == 0) buf[(size_t)blockIdx.x * NV + k] = x;
        }
    }
    grid.sync();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(&buf[(size_t)b * NV + k]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(kFull, s, off);
            if (lane == 0) s_bc[k] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = s_bc[k];
}

// ---- several ranks: the two sums of an iteration also run over the NVLink mailboxes ----------------
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
constexpr long long kPeerWait = 6000000000LL;    // ~3 s: a lost peer must not hang the GPU
constexpr long long kLocalWait = 30000000000LL;  // block 0 always answers within kPeerWait

// Two 8-byte words per double, each carrying 32 payload bits and the low 32 bits of the sequence number
// (the self-validating words of kernels.cu's p2p_allreduce).
__device__ __forceinline__ void put_tagged(unsigned long long *dst, double v, unsigned long long tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    st_relaxed_sys(&dst[0], (bits & 0xffffffffull) | tag);
    st_relaxed_sys(&dst[1], (bits >> 32) | tag);
}
__device__ __forceinline__ bool get_tagged(const unsigned long long *src, unsigned long long tag, long long limit,
                                           double &out) {
    const long long t0 = clock64();
    unsigned long long lo, hi;
    for (;;) {
        lo = ld_relaxed_sys(&src[0]);
        hi = ld_relaxed_sys(&src[1]);
        if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
        if (clock64() - t0 > limit) return false;
    }
    out = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return true;
}

// On entry v[] holds this rank's totals (identical in every thread of the grid); on return the totals
// over all ranks, identical in every thread of every rank.  Block 0 is the only block that talks to the
// peers: it stores its words into every rank's mailbox, collects the ranks' words from its own mailbox
// (bounded wait), adds them in rank order and publishes the result -- or NaN after a timeout -- in the
// local `pub` words, which every block reads.  One source of truth: all blocks take the same branch.
template <int NV>
__device__ __forceinline__ void peer_allsum(const PeerComm &pc, unsigned long long *pub, unsigned long long seq,
                                            double (&v)[NV], double *s_bc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = (int)(seq & 1ull);
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    unsigned long long *mine = pub + slot * 4;
    __syncthreads();  // every thread has taken its copy of the previous reduction out of s_bc
    if (warp == 0) {
        if (blockIdx.x == 0) {
            double got[NV];
            bool ok = true;
#pragma unroll
            for (int k = 0; k < NV; ++k) got[k] = 0.0;
            if (lane < pc.nranks) {
#pragma unroll
                for (int k = 0; k < NV; ++k) put_tagged(&pc.box[lane]->ll[slot][pc.rank][2 * k], v[k], tag);
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    ok = get_tagged(&pc.box[pc.rank]->ll[slot][lane][2 * k], tag, kPeerWait, got[k]) && ok;
            }
            ok = __all_sync(kFull, ok);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double t = 0.0;
                for (int r = 0; r < pc.nranks; ++r) t += __shfl_sync(kFull, got[k], r);  // rank order everywhere
                if (!ok) t = __longlong_as_double(0x7ff8000000000000LL);
                if (lane == 0) put_tagged(&mine[2 * k], t, tag);
            }
            if (!ok && lane == 0) *pc.status = 1;
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double t;
                if (!get_tagged(&mine[2 * k], tag, kLocalWait, t)) t = __longlong_as_double(0x7ff8000000000000LL);
                s_bc[k] = t;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = s_bc[k];
    __syncthreads();  // s_bc is written again by the next reduction of this block
}

// one row of A times the staged part of d: ascending columns, separate multiply and add (the serial
// CSR arithmetic of k_spmv).  Padding entries (value 0) may carry a column outside the staged range:
// the index is clamped into it.
template <int CH>
__device__ __forceinline__ double row_times_staged(const Sell &A, const double *__restrict__ val, uint32_t base,
                                                   int len, const double *sd, int c0, unsigned cn) {
    double s = 0.0;
    for (int k0 = 0; k0 < len; k0 += CH) {
        int c[CH];
        double v[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int kk = min(k0 + k, len - 1);
            const uint32_t q = base + (uint32_t)kk * kSlice;
            c[k] = A.col[q];
            v[k] = val[q];
        }
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (k0 + k < len) {  // warp-uniform
                const unsigned idx = min((unsigned)(c[k] - c0), cn - 1u);
                s = __dadd_rn(s, __dmul_rn(v[k], sd[idx]));
            }
    }
    return s;
}

template <int CH, bool PEERS>
__global__ void __launch_bounds__(kFusedThreads, 1) k_cg_fused(CgFusedArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];
    __shared__ double s_red[2 * 32];
    __shared__ double s_bc[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *sx = smem;                                         // x of my rows   [wpb][1024]
    double *sg = smem + (size_t)a.wpb * kFusedThreads;         // g of my rows   [wpb][1024]
    double *sd = smem + 2 * (size_t)a.wpb * kFusedThreads;     // d of the block's column range
    const int c0 = a.blk_c0[blockIdx.x];
    const unsigned cn = (unsigned)a.blk_cn[blockIdx.x];
    const int diag0 = a.own_off - c0;                          // staged index of row r's own column = diag0 + r

    // scalars of the recurrences: carried redundantly, and identically, by every thread of the grid
    double gh_old = a.S->gh[0], gh_new = a.S->gh_new, gg = a.S->gg, res = a.S->res, dAd = 0.0;
    const double reduced_tol = a.S->reduced_tol, tol = a.S->tol;
    const int maxit = a.S->maxit;
    int it = a.S->it, status = a.S->status[0];
    if (status != 0) return;  // the start residual already met the criterion (same answer in every block)

    int rk[kFusedMaxWin];
    double hk[kFusedMaxWin];
#pragma unroll
    for (int k = 0; k < kFusedMaxWin; ++k) {
        rk[k] = -1;
        hk[k] = 0.0;
        const int win = blockIdx.x * a.wpb + k;
        if (k < a.wpb && win < a.nwin) {
            const int r = a.A.row_of[(size_t)win * kWindow + tid];
            rk[k] = r;
            if (r >= 0) {
                sx[k * kFusedThreads + tid] = a.x_own[r];
                sg[k * kFusedThreads + tid] = a.g[r];
            }
        }
    }

    const size_t pstride = (size_t)gridDim.x * 2;
    // several ranks: does this block's staged range reach into the ghost blocks of d?
    const bool reads_lo_ghost = PEERS && a.pc.rank > 0 && c0 < a.own_off;
    const bool reads_hi_ghost = PEERS && a.pc.rank < a.pc.nranks - 1 && c0 + (int)cn > a.own_off + a.A.nrows;
    const int hi0 = a.A.nrows - a.pc.hi_count;
    for (int iter = 0;; ++iter) {
        if (PEERS && iter > 0 && (reads_lo_ghost || reads_hi_ghost)) {
            // the neighbours stored their boundary blocks of d into my ghost blocks in their phase 3 and
            // raised the flag after their grid barrier (iteration 0: the host exchanged the halo)
            if (tid == 0) {
                const unsigned long long want = a.halo_seq0 + (unsigned long long)iter;
                const PeerMailbox *box = a.pc.box[a.pc.rank];
                const long long t0 = clock64();
                bool ok = true;
                while (ok && reads_lo_ghost && ld_acquire_sys(&box->halo_flag[0]) < want) ok = clock64() - t0 < kPeerWait;
                while (ok && reads_hi_ghost && ld_acquire_sys(&box->halo_flag[1]) < want) ok = clock64() - t0 < kPeerWait;
                if (!ok) *a.pc.status = 1;  // the sums stay global, so every block still takes the same branches
            }
            __syncthreads();
        }
        // the block's part of d (written by all blocks in phase 3 of the previous iteration, or by the
        // start kernel): L2 loads, this SM's L1 may hold last iteration's lines
        for (unsigned i = tid; i < cn; i += kFusedThreads) sd[i] = __ldcg(&a.d[c0 + i]);
        __syncthreads();

        // phase 1: h = A d, dAd = d . h
        double acc1[1] = {0.0};
#pragma unroll
        for (int k = 0; k < kFusedMaxWin; ++k) {
            const int win = blockIdx.x * a.wpb + k;
            if (k < a.wpb && win < a.nwin) {
                const int slice = win * (kWindow / kSlice) + warp;
                const uint32_t b0 = a.A.slice_ptr[slice];
                const int len = (int)((a.A.slice_ptr[slice + 1] - b0) >> 5);
                const double s = row_times_staged<CH>(a.A, a.val, b0 + lane, len, sd, c0, cn);
                hk[k] = s;
                if (rk[k] >= 0) acc1[0] += s * sd[diag0 + rk[k]];
            }
        }
        grid_allsum<1>(grid, acc1, a.partials, s_red, s_bc);
        if (PEERS) peer_allsum<1>(a.pc, a.pub, a.ar_seq0 + 2ull * iter + 1ull, acc1, s_bc);
        dAd = acc1[0];
        const double alpha = gh_old / dAd;

        // phase 2: g += alpha h ; gg = g . g ; h = D^-1 g ; gh' = g . h
        double acc2[2] = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < kFusedMaxWin; ++k)
            if (rk[k] >= 0) {
                const double gi = sg[k * kFusedThreads + tid] + alpha * hk[k];
                sg[k * kFusedThreads + tid] = gi;
                acc2[0] += gi * gi;
                const double hi = a.dinv[rk[k]] * gi;
                hk[k] = hi;
                acc2[1] += gi * hi;
            }
        grid_allsum<2>(grid, acc2, a.partials + pstride, s_red, s_bc);
        if (PEERS) peer_allsum<2>(a.pc, a.pub, a.ar_seq0 + 2ull * iter + 2ull, acc2, s_bc);
        gg = acc2[0];
        gh_new = acc2[1];

        // iteration_status(it, res) of ReductionControl(maxit, tol, reduce), then beta
        res = sqrt(fabs(gg));
        ++it;
        if (res <= reduced_tol || res <= tol) status = 1;
        else if (it >= maxit || isnan(res)) status = 2;
        const double beta = gh_new / gh_old;

        // phase 3: x += alpha d ; d = beta d - h (published for the next SpMV of every block)
#pragma unroll
        for (int k = 0; k < kFusedMaxWin; ++k)
            if (rk[k] >= 0) {
                const double dold = sd[diag0 + rk[k]];
                sx[k * kFusedThreads + tid] += alpha * dold;
                if (status == 0) {
                    const double dnew = beta * dold - hk[k];
                    a.d[a.own_off + rk[k]] = dnew;
                    if (PEERS) {  // my first / last block of rows is the neighbours' ghost block
                        if (a.pc.d_lo && rk[k] < a.pc.lo_count) a.pc.d_lo[rk[k]] = dnew;
                        if (a.pc.d_hi && rk[k] >= hi0) a.pc.d_hi[rk[k] - hi0] = dnew;
                    }
                }
            }
        gh_old = gh_new;
        if (status != 0) break;  // same decision in every thread of the grid (and of every rank)
        if (PEERS) __threadfence_system();  // my halo stores are ordered before the barrier and the flag
        grid.sync();
        if (PEERS && blockIdx.x == 0 && tid == 0) {
            const unsigned long long seq = a.halo_seq0 + (unsigned long long)iter + 1ull;
            __threadfence_system();
            if (a.pc.rank > 0) st_release_sys(&a.pc.box[a.pc.rank - 1]->halo_flag[1], seq);
            if (a.pc.rank < a.pc.nranks - 1) st_release_sys(&a.pc.box[a.pc.rank + 1]->halo_flag[0], seq);
        }
    }

#pragma unroll
    for (int k = 0; k < kFusedMaxWin; ++k)
        if (rk[k] >= 0) a.x_own[rk[k]] = sx[k * kFusedThreads + tid];
    if (blockIdx.x == 0 && tid == 0) {
        a.S->it = it;
        a.S->res = res;
        a.S->status[0] = status;  // a peer wait that timed out is recorded in CgScalars::peer_timeout
        a.S->gg = gg;
        a.S->gh_new = gh_new;
        a.S->gh[0] = gh_old;
        a.S->dAd = dAd;
    }
}

__global__ void __launch_bounds__(kWindow) k_window_col_range(Sell A, int nwin, int32_t *cmin, int32_t *cmax) {
    __shared__ int s_lo[kWindow / 32], s_hi[kWindow / 32];
    const int win = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lo = INT_MAX, hi = -1;
    if (win < nwin) {
        const size_t slot = (size_t)win * kWindow + tid;
        const int r = A.row_of[slot];
        if (r >= 0) {
            const uint32_t base = A.slice_ptr[slot >> 5] + (uint32_t)lane;
            const uint32_t len = A.rowptr[r + 1] - A.rowptr[r];
            for (uint32_t k = 0; k < len; ++k) {
                const int c = A.col[base + kSlice * k];
                lo = min(lo, c);
                hi = max(hi, c);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(kFull, lo, off));
        hi = max(hi, __shfl_xor_sync(kFull, hi, off));
    }
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
    if (warp == 0) {
        lo = s_lo[lane];
        hi = s_hi[lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo = min(lo, __shfl_xor_sync(kFull, lo, off));
            hi = max(hi, __shfl_xor_sync(kFull, hi, off));
        }
        if (lane == 0 && win < nwin) { cmin[win] = lo; cmax[win] = hi; }
    }
}

void note(const Launcher &l, cudaError_t e) {
    if (l.count) ++*l.count;
    if (l.error && *l.error == cudaSuccess) *l.error = e != cudaSuccess ? e : cudaPeekAtLastError();
}

using FusedKernel = void (*)(CgFusedArgs);
FusedKernel kernel_for(const Sell &A, bool peers) {
    if (peers) return A.chunk <= 7 ? k_cg_fused<7, true> : k_cg_fused<10, true>;
    return A.chunk <= 7 ? k_cg_fused<7, false> : k_cg_fused<10, false>;
}

}  // namespace

void launch_window_col_range(const Launcher &l, const Sell &A, int nwin, int32_t *cmin, int32_t *cmax) {
    k_window_col_range<<<nwin, kWindow, 0, l.stream>>>(A, nwin, cmin, cmax);
    note(l, cudaSuccess);
}

size_t cg_fused_smem_bytes(int wpb, int stage_cap) {
    return sizeof(double) * (2 * (size_t)wpb * kFusedThreads + (size_t)stage_cap);
}

bool cg_fused_supported(int grid, size_t smem_bytes) {
    int dev = 0, coop = 0, sms = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (!coop || smem_bytes + 1024 > (size_t)optin) return false;  // 1 KiB left for the static arrays
    for (FusedKernel k : {FusedKernel(k_cg_fused<7, false>), FusedKernel(k_cg_fused<10, false>),
                          FusedKernel(k_cg_fused<7, true>), FusedKernel(k_cg_fused<10, true>)}) {
        // always the device maximum: contexts with different plans share the kernel's attribute
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kFusedThreads, smem_bytes) != cudaSuccess ||
            (int64_t)occ * sms < grid) {
            cudaGetLastError();
            return false;
        }
    }
    return true;
}

cudaError_t launch_cg_fused(const Launcher &l, int grid, size_t smem_bytes, const CgFusedArgs &a) {
    CgFusedArgs args = a;
    void *params[] = {&args};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)kernel_for(a.A, a.pc.enabled != 0), dim3((unsigned)grid),
                                                      dim3(kFusedThreads), params, smem_bytes, l.stream);
    note(l, e);
    return e;
}

}  // namespace wv
