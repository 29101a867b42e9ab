// kernels_peer.cuh — the exchange step of the sharded full-energy evaluation over NVLink peer memory.
//
// the only multi-GPU exchange on the path is the sum of each rank's partial vector
// (8 scalars + NK complex ρ(k) partials = 5.5 KB at nk = 5).  A library all-reduce of that payload is pure
// latency (two launches, a ring/tree protocol and its synchronisation: ≈100 µs per evaluation measured through
// torch.distributed at 8 GPUs, a third of the whole step).  Here every rank owns an exchange buffer that its
// peers map through CUDA IPC:
//     k_peer_push  rank r stores its vector into slot r of EVERY rank's buffer (plain st.global to peer
//                  addresses: the data crosses NVLink/NVSwitch once per peer), __threadfence_system, then an
//                  epoch flag per destination;
//     k_peer_sum   waits for the `world` flags of this epoch in its OWN memory and adds the slots in rank order
//                  (the same order on every rank: bit-identical totals everywhere), then the evaluation's own
//                  finalisation (E_recip, Properties) runs on the summed vector as before.
// Buffers are double-buffered by epoch parity: a rank can only reach epoch e+2 after every peer pushed e+1,
// which each peer does after it finished reading e — so one flag wait per evaluation is the whole protocol.
#pragma once
#include "mmc_common.cuh"

#define MMC_PEER_MAX 8

struct PeerArgs {
    double *slot[MMC_PEER_MAX];                 // exchange buffer of rank d (peer-mapped), [2][world][nvec_cap]
    unsigned long long *flag[MMC_PEER_MAX];     // flags of rank d, [2][world]
    int world, rank, nvec, nvec_cap, parity;
    unsigned long long epoch;
};

static __global__ void __launch_bounds__(256) k_peer_push(const __grid_constant__ PeerArgs P, const double *__restrict__ vec)
{
    const int d = blockIdx.x;                   // destination rank
    double *dst = P.slot[d] + ((size_t)P.parity * P.world + P.rank) * P.nvec_cap;
    for (int t = threadIdx.x; t < P.nvec; t += 256) dst[t] = vec[t];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned long long *f = P.flag[d] + (size_t)P.parity * P.world + P.rank;
        *f = P.epoch;
    }
}

// out[0 .. nvec): Σ_ranks slot; status[0] = 1 when a peer's flag did not arrive within the spin limit
static __global__ void __launch_bounds__(256) k_peer_sum(const __grid_constant__ PeerArgs P, double *__restrict__ out, int *status)
{
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < P.world) {
        volatile unsigned long long *f = P.flag[P.rank] + (size_t)P.parity * P.world + threadIdx.x;
        const long long t0 = clock64();
        while (*f < P.epoch) {
            __nanosleep(20);
            if (clock64() - t0 > 20000000000LL) { s_bad = 1; break; }     // ≈10 s: a peer is gone
        }
    }
    __syncthreads();
    __threadfence_system();
    const volatile double *base = P.slot[P.rank] + (size_t)P.parity * P.world * P.nvec_cap;
    for (int t = threadIdx.x; t < P.nvec; t += 256) {
        double s = 0.0;
        for (int q = 0; q < P.world; ++q) s += base[(size_t)q * P.nvec_cap + t];
        out[t] = s;
    }
    if (threadIdx.x == 0 && s_bad) *status = 1;
}

// k_peer_sum with the evaluation's own finalisation folded in (one launch after the exchange): the summed vector, then
// E_recip = Σ cfac |ρ(k)|² (ewalds.jl:599), ρ(k) into the resident buffers (:600-601) and the scalars into the mapped host slot.
struct PeerFinishArgs {
    int nkvecs;                    // 0: no k-space
    const double *cfac; double2 *dst0, *dst1;
    double *host_out;              // mapped pinned: [0..7] summed scalars, [8] sequence number (written last), [9] status
    unsigned long long seq;
};

// (body shared with k_eval_tail, whose last block runs it right after its own push: one launch less per sharded evaluation)
__device__ __forceinline__ void peer_sum_finish_block(const PeerArgs &P, double *__restrict__ out, const PeerFinishArgs &F)
{
    __shared__ int s_bad;
    __shared__ double s_red[8];
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < P.world) {
        volatile unsigned long long *f = P.flag[P.rank] + (size_t)P.parity * P.world + threadIdx.x;
        const long long t0 = clock64();
        while (*f < P.epoch) {
            __nanosleep(20);
            if (clock64() - t0 > 20000000000LL) { s_bad = 1; break; }     // ≈10 s: a peer is gone
        }
    }
    __syncthreads();
    __threadfence_system();
    const volatile double *base = P.slot[P.rank] + (size_t)P.parity * P.world * P.nvec_cap;
    double er[1] = {0.0};
    for (int t = threadIdx.x; t < P.nvec; t += 256) {
        double s = 0.0;
        for (int q = 0; q < P.world; ++q) s += base[(size_t)q * P.nvec_cap + t];
        out[t] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < F.nkvecs; k += 256) {
        const double2 s = make_double2(out[MMC_NSCAL + 2 * k], out[MMC_NSCAL + 2 * k + 1]);
        er[0] += F.cfac[k] * (s.x * s.x + s.y * s.y);
        if (F.dst0) F.dst0[k] = s;
        if (F.dst1) F.dst1[k] = s;
    }
    block_sum<1, 256>(er, s_red);
    if (threadIdx.x == 0) {
        out[4] = er[0];
        for (int i = 0; i < MMC_NSCAL; ++i) F.host_out[i] = out[i];
        F.host_out[MMC_NSCAL + 1] = s_bad ? 1.0 : 0.0;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(F.host_out + MMC_NSCAL) = F.seq;
    }
}

static __global__ void __launch_bounds__(256) k_peer_sum_finish(const __grid_constant__ PeerArgs P, double *__restrict__ out,
                                                                 const __grid_constant__ PeerFinishArgs F)
{
    peer_sum_finish_block(P, out, F);
}
