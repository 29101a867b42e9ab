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

// ------------------------------------------------------------------------------------------------------------
// COM all-gather for the domain-decomposed host evaluation: every rank needs ALL centres of mass to bin and partition
// identically, but only 1/world of them has to cross ITS PCIe link — rank r copies its slice of the caller's COM array into
// the staging area of its own exchange buffer, publishes "slice r of call e is there" into every rank's flag array
// (k_com_publish, after the copy in stream order), and k_repack_com_gather builds the resident COM array reading slice q
// from rank q's staging area over NVLink (plain loads from peer memory) once flag q shows call e.  A rank overwrites its
// slice for call e + 1 only after it has the result of call e, which needs every rank's partial vector of call e, which every
// rank pushes after its own gather kernel — so no reader can still be on the old slice: no double buffering.
// ------------------------------------------------------------------------------------------------------------
struct ComGatherArgs {
    const double *stage[MMC_PEER_MAX];        // raw [3 x n_mol] staging areas of all ranks (only slice q of stage[q] is current)
    unsigned long long *flag[MMC_PEER_MAX];   // flag[q] + r: rank q's copy of "rank r's slice is in place" (epoch)
    unsigned long long epoch;
    int world, rank, n_mol;
    double box;
    double4 *dcom;
    int *info;
    unsigned int *zero_words; int n_zero_words;    // k_com_wait clears these (the block-need flags)
    int *zero_count; int n_zero_count;             // k_repack_com_gather clears these (the cell populations, before k_bin7)
};

// slices begin on multiples of 256 molecules: a block of the gather kernel has one source rank and its 6 KB chunk is 16-byte aligned
__host__ __device__ __forceinline__ int com_slice_begin(int n_mol, int world, int q)
{
    return q >= world ? n_mol : (int)((long long)n_mol * q / world) / 256 * 256;
}

// (the slice was written by a host->device copy that completed earlier in this stream: nothing of this kernel's to fence)
static __global__ void k_com_publish(ComGatherArgs A)
{
    if ((int)threadIdx.x < A.world) {
        volatile unsigned long long *f = A.flag[threadIdx.x] + A.rank;
        *f = A.epoch;
    }
}

// waits (one thread per rank) until every rank's slice of this call is in place; the gather kernel behind it in the stream needs
// no flag and no fence of its own (with a system-scope fence in each of its 1000 blocks it took 140 us whenever this GPU's PCIe
// link was busy with the site copies — the fence's round trip queues behind the DMA traffic — against 25 us on an idle link)
static __global__ void k_com_wait(ComGatherArgs A)
{
    // (also clears what the kernels behind it accumulate into: the validation word and the block-need flags — two memsets less)
    if (threadIdx.x < 4) A.info[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < A.n_zero_words; i += blockDim.x) A.zero_words[i] = 0u;
    __syncthreads();
    if ((int)threadIdx.x < A.world) {
        volatile unsigned long long *f = A.flag[A.rank] + threadIdx.x;
        const long long c0 = clock64();
        while (*f < A.epoch) {
            __nanosleep(20);
            if (clock64() - c0 > 20000000000LL) { atomicOr(&A.info[0], REPACK_PEER_TIMEOUT); break; }     // ≈10 s: a peer is gone
        }
    }
    __threadfence_system();
}

// one block = 256 molecules = 768 doubles of the raw array: read as 384 16-byte loads from the owner's staging area (peer memory:
// 8-byte loads at a 24-byte stride cost three sectors per sector used), turned through shared memory
static __global__ void __launch_bounds__(256) k_repack_com_gather(const __grid_constant__ ComGatherArgs A)
{
    __shared__ double2 s_raw[384];
    const int t0 = blockIdx.x * 256, t = t0 + threadIdx.x;
    for (int i = t; i < A.n_zero_count; i += gridDim.x * 256) A.zero_count[i] = 0;
    int q = (int)((long long)t0 * A.world / A.n_mol);
    while (q > 0 && t0 < com_slice_begin(A.n_mol, A.world, q)) --q;
    while (q + 1 < A.world && t0 >= com_slice_begin(A.n_mol, A.world, q + 1)) ++q;
    const double *src = A.stage[q] + 3 * (size_t)t0;
    const int nd = 3 * min(256, A.n_mol - t0);            // doubles of this block's chunk
    for (int i = threadIdx.x; 2 * i < nd; i += 256) {
        if (2 * i + 1 < nd) s_raw[i] = __ldcg(reinterpret_cast<const double2 *>(src) + i);
        else s_raw[i] = make_double2(__ldcg(src + 2 * i), 0.0);
    }
    __syncthreads();
    if (t >= A.n_mol) return;
    const double *r = reinterpret_cast<const double *>(s_raw) + 3 * threadIdx.x;
    const double x = r[0], y = r[1], z = r[2];
    if (!(x >= 0.0 && x <= A.box && y >= 0.0 && y <= A.box && z >= 0.0 && z <= A.box)) atomicOr(&A.info[0], REPACK_COM_OUTSIDE);
    A.dcom[t] = make_double4(x, y, z, 0.0);
}

// small results to the host through MAPPED pinned memory instead of a device->host copy: a copy would queue in the copy engine
// behind the host->device site copies another stream has in flight (mmc_potential_host), and the host waits for exactly these bytes
static __global__ void k_bytes_to_host(const unsigned char *__restrict__ src, int n, const int *__restrict__ info, int n_info,
                                       unsigned char *h_dst, int *h_info)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) h_dst[i] = src[i];
    if ((int)threadIdx.x < n_info) h_info[threadIdx.x] = info[threadIdx.x];
    __threadfence_system();
}
