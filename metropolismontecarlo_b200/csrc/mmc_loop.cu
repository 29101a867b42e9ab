// mmc_loop.cu — host-side stand-in for the Julia move loop, built ON TOP of the C ABI.
//
// The north star keeps Loop() (Ewald/main.jl:460-696) in Julia; Julia is not installed where this
// library is built and tested, so this file plays Julia's part in C++: it generates the trial
// moves exactly as the reference does, calls mmc_trial_move / mmc_accept / mmc_reject once per
// move (one kernel launch + one host wait each, what a ccall-driven loop would pay) and applies
// the Metropolis rule.  Random numbers come from a caller-supplied stream of uniforms consumed
// in the reference's draw order, so a recorded Julia stream reproduces the
// reference trajectory.  All energies come from the CUDA kernels; nothing here evaluates one.

#include "mmc_handle.h"
#include "julia_mt.h"
#include "kernels_chain.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace mmc_detail;

#include <map>
#include <mutex>
namespace {
// cudaFuncSetAttribute on a kernel that has launches in flight makes the driver wait for them: with one call per block of
// moves, chains of different handles (streams) ran one after the other (measured: 16 replicas on one GPU at 2x the rate of one).
// The opt-in shared-memory size is therefore raised once per kernel and only when a launch needs more than what is set.
cudaError_t ensure_smem_optin(const void *func, int bytes)
{
    static std::mutex mtx;
    static std::map<const void *, int> set;
    std::lock_guard<std::mutex> g(mtx);
    auto it = set.find(func);
    if (it != set.end() && it->second >= bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) set[func] = bytes;
    return e;
}
}  // namespace

namespace {

struct UStream {
    const double *u; int64_t n, pos; bool dry;
    // past the end of the caller's stream: `dry` is raised and the move in progress is abandoned by the loops below —
    // it is rejected without being recorded, its counters are taken back and uniforms_used returns to the start of the
    // move, so that a caller who refills the stream resumes exactly there (the returned 0.5 is never acted upon)
    double next() { if (pos >= n) { dry = true; return 0.5; } return u[pos++]; }
};

// Ewald/boundaries.jl:16-26
inline void pbc3(double v[3], double box)
{
    for (int k = 0; k < 3; ++k) {
        if (v[k] > box) v[k] -= box;
        if (v[k] < 0) v[k] += box;
    }
}

// Ewald/auxillary.jl:106-114 — a uniform is drawn only when delta >= 0
inline bool metropolis(double delta, UStream &us)
{
    if (delta < 0.0) return true;
    return std::exp(-delta) > us.next();
}

// Ewald/adjust.jl:1-83 (Adjust! and Adjust_rot! share one body)
struct MoveStat { int64_t naccepp = 0, naccept = 0, attempp = 0, attempt = 0; double set_value = 0.5, d_max = 0; };
inline void adjust_step(MoveStat &m, double L)
{
    if (m.attempp == 0) { m.naccepp = m.naccept; m.attempp = m.attempt; return; }
    const double ratio = double(m.naccept - m.naccepp) / double(m.attempt - m.attempp);
    const double old = m.d_max;
    m.d_max = m.d_max * ratio / m.set_value;
    const double r = m.d_max / old;
    if (r > 1.5) m.d_max = old * 1.5;
    if (r < 0.5) m.d_max = old * 0.5;
    if (m.d_max > L / 2) m.d_max = L / 2;
    m.naccepp = m.naccept; m.attempp = m.attempt;
}

// Ewald/quaternions.jl:37-50 — rows as written in the reference, [2,3] = 2(q2 q4 + q1 q2)
inline void quat_to_matrix(const double q[4], double a[3][3])
{
    const double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
    a[0][0] = q1 * q1 + q2 * q2 - q3 * q3 - q4 * q4; a[0][1] = 2 * (q2 * q3 + q1 * q4); a[0][2] = 2 * (q2 * q4 - q1 * q3);
    a[1][0] = 2 * (q2 * q3 - q1 * q4); a[1][1] = q1 * q1 - q2 * q2 + q3 * q3 - q4 * q4; a[1][2] = 2 * (q2 * q4 + q1 * q2);
    a[2][0] = 2 * (q2 * q4 + q1 * q3); a[2][1] = 2 * (q3 * q4 - q1 * q2); a[2][2] = q1 * q1 - q2 * q2 - q3 * q3 + q4 * q4;
}

}  // namespace

extern "C" int mmc_loop_run(mmc_handle *h, const mmc_loop_params *p, double *com, double *quat, const double *db,
                            const double *uniforms, int64_t n_uniforms, int64_t n_moves, double e0, double v0,
                            uint8_t *accepted, double *delta_out, mmc_loop_stats *st)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system) FAIL(MMC_ESTATE, "no molecular system uploaded");
    if (!p || !com || !quat || !db || !uniforms || !st) FAIL(MMC_EINVAL, "null argument");
    int rc = style_check(h, p->style);
    if (rc) return rc;
    const double box = h->S.box;
    if (!(h->S.rc_lj < box / 2)) FAIL(MMC_EINVAL, "r_cut must be < box/2 (Ewald/main.jl:483)");
    UStream us{uniforms, n_uniforms, 0, false};
    MoveStat tr, ro;
    tr.d_max = p->dr_max; ro.d_max = p->dphi_max;
    double dr_max = p->dr_max, dphi_max = p->dphi_max;
    std::memset(st, 0, sizeof(*st));
    st->total_energy = e0; st->total_virial = v0;
    const int n_mol = h->S.n_mol;
    double sites[3 * MMC_MAX_SITES];
    int ret = 0;
    for (int64_t m = 0; m < n_moves; ++m) {
        const int i = (int)(m % n_mol);                  // sweep order i = 1..N (main.jl:490)
        const int2 mi = mol_of(h, i);
        const int64_t pos_m = us.pos;                    // resume point if the stream runs dry inside this move
        double rnew[3] = {com[3 * i], com[3 * i + 1], com[3 * i + 2]};
        double ei[4];
        const double chose_move = us.next();             // main.jl:516
        bool is_trans;
        if (chose_move < p->p_trans) {                   // main.jl:519-529, auxillary.jl:94-103
            is_trans = true;
            tr.attempt += 1;
            const double z0 = us.next(), z1 = us.next(), z2 = us.next();
            rnew[0] = rnew[0] + (z0 - 0.5) * dr_max;
            rnew[1] = rnew[1] + (z1 - 0.5) * dr_max;
            rnew[2] = rnew[2] + (z2 - 0.5) * dr_max;
            pbc3(rnew, box);
            std::memcpy(ei, quat + 4 * i, sizeof(ei));
        } else if (chose_move <= p->p_rot) {             // main.jl:530-538, quaternions.jl:52-73,94-120,158-182
            is_trans = false;
            ro.attempt += 1;
            const double *old = quat + 4 * i;
            if (std::fabs(old[0] * old[0] + old[1] * old[1] + old[2] * old[2] + old[3] * old[3] - 1.0) > 1.e-6) { ret = 2; break; }
            double ax[3], nrm;
            for (;;) {
                ax[0] = 2.0 * us.next() - 1.0; ax[1] = 2.0 * us.next() - 1.0; ax[2] = 2.0 * us.next() - 1.0;
                nrm = ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2];
                if (nrm < 1.0 || us.dry) break;
            }
            const double sn = std::sqrt(nrm);
            ax[0] = ax[0] / sn; ax[1] = ax[1] / sn; ax[2] = ax[2] / sn;
            const double zeta = us.next();
            const double angle = (2.0 * zeta - 1.0) * dphi_max;
            const double rq[4] = {std::cos(0.5 * angle), std::sin(0.5 * angle) * ax[0], std::sin(0.5 * angle) * ax[1],
                                  std::sin(0.5 * angle) * ax[2]};
            ei[0] = rq[0] * old[0] - rq[1] * old[1] - rq[2] * old[2] - rq[3] * old[3];   // quatmul(rot, old)
            ei[1] = rq[1] * old[0] + rq[0] * old[1] - rq[3] * old[2] + rq[2] * old[3];
            ei[2] = rq[2] * old[0] + rq[3] * old[1] + rq[0] * old[2] - rq[1] * old[3];
            ei[3] = rq[3] * old[0] - rq[2] * old[1] + rq[1] * old[2] + rq[0] * old[3];
        } else { ret = 3; break; }                       // main.jl:539-541
        if (us.dry) {                                    // not enough uniforms to draw this move: it never happened
            if (is_trans) tr.attempt -= 1; else ro.attempt -= 1;
            us.pos = pos_m; ret = 1; break;
        }
        if (std::fabs(ei[0] * ei[0] + ei[1] * ei[1] + ei[2] * ei[2] + ei[3] * ei[3] - 1.0) > 1.e-6) { ret = 2; break; }
        double a[3][3];
        quat_to_matrix(ei, a);
        for (int s = 0; s < mi.y; ++s) {                 // main.jl:545-548: COM + MATMUL(ai, db)
            const double *d = db + 3 * (mi.x + s);
            for (int c = 0; c < 3; ++c)
                sites[3 * s + c] = rnew[c] + (d[0] * a[0][c] + d[1] * a[1][c] + d[2] * a[2][c]);
        }
        mmc_trial_result r;
        if ((rc = mmc_trial_move(h, i + 1, rnew, sites, p->style, &r))) return rc;
        double partial_old_e = r.lj_old, partial_old_v = r.lj_vir_old;
        double partial_new_e = r.lj_new, partial_new_v = r.lj_vir_new;
        if (p->style != MMC_STYLE_LJ_ONLY) {             // main.jl:501-505, 566-570
            partial_old_v += r.qq_vir_old; partial_old_e += r.qq_old;
            partial_new_v += r.qq_vir_new; partial_new_e += r.qq_new;
        }
        const bool overlap = r.overlap_old || r.overlap_new;
        const double deltaRecip = r.d_recip;             // already 0 on overlap / non-Ewald
        const double delta = (partial_new_e) - (partial_old_e) + deltaRecip;   // main.jl:593
        const bool acc = metropolis(delta / p->temperature, us) && !overlap;   // main.jl:598
        if (us.dry) {                                    // Metropolis found the stream empty: reject, record nothing, resume at pos_m
            if ((rc = mmc_reject(h))) return rc;
            if (is_trans) tr.attempt -= 1; else ro.attempt -= 1;
            us.pos = pos_m; ret = 1; break;
        }
        if (overlap) st->n_overlap += 1;
        if (acc) {
            st->total_energy += delta;
            st->total_virial += (partial_new_v - partial_old_v) + deltaRecip / 3;
            st->n_accepted += 1;
            if (is_trans) tr.naccept += 1; else ro.naccept += 1;
            com[3 * i] = rnew[0]; com[3 * i + 1] = rnew[1]; com[3 * i + 2] = rnew[2];
            std::memcpy(quat + 4 * i, ei, sizeof(ei));
            if ((rc = mmc_accept(h))) return rc;
        } else {
            if ((rc = mmc_reject(h))) return rc;
        }
        if (accepted) accepted[m] = acc ? 1 : 0;
        if (delta_out) delta_out[m] = delta;
        if (p->adjust && i == n_mol - 1) {               // main.jl:645-651
            tr.d_max = dr_max; adjust_step(tr, box); dr_max = tr.d_max;
            ro.d_max = dphi_max; adjust_step(ro, box); dphi_max = ro.d_max;
        }
        st->n_moves = m + 1;
    }
    st->uniforms_used = us.pos;
    st->trans_attempt = tr.attempt; st->trans_accept = tr.naccept;
    st->rot_attempt = ro.attempt; st->rot_accept = ro.naccept;
    st->dr_max = dr_max; st->dphi_max = dphi_max;
    return ret;
}

// The same loop, evaluated on the device for the whole block of moves (kernels_chain.cuh): one launch,
// the system's state in one SM's shared memory.  Arguments, draw order, return codes and statistics are
// those of mmc_loop_run; energies are summed in a different (fixed) order than the per-move kernels, so
// deltas agree to rounding (≈1e-13 relative) and the accept/reject record is the same.
extern "C" int mmc_loop_run_device(mmc_handle *h, const mmc_loop_params *p, double *com, double *quat, const double *db,
                                   const double *uniforms, int64_t n_uniforms, int64_t n_moves, double e0, double v0,
                                   uint8_t *accepted, double *delta_out, mmc_loop_stats *st)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system) FAIL(MMC_ESTATE, "no molecular system uploaded");
    if (!p || !com || !quat || !db || !uniforms || !st || n_moves < 0 || n_uniforms < 0) FAIL(MMC_EINVAL, "bad argument");
    int rc = style_check(h, p->style);
    if (rc) return rc;
    if (p->style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "mmc_loop_run_device is for molecular systems");
    if (h->trial_pending || h->vol_pending) FAIL(MMC_ESTATE, "a trial move is pending");
    const DevSystem &S = h->S;
    if (!(S.rc_lj < S.box / 2)) FAIL(MMC_EINVAL, "r_cut must be < box/2 (Ewald/main.jl:483)");
    if (!h->uniform || h->US < 1 || h->US > 4) FAIL(MMC_EINVAL, "device loop needs a uniform topology with 1..4 sites per molecule");
    if (p->style == MMC_STYLE_EWALD && S.nk > MMC_MAX_NK) FAIL(MMC_EINVAL, "the per-move k-space kernels are sized for nk <= 8");
    const bool recip = p->style == MMC_STYLE_EWALD;
    int dev = 0, max_optin = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // thread-block cluster version (k_chains): C CTAs on C SMs, each holding a slice of the molecules and of the k-vectors;
    // one CTA (k_chain) holds the whole system
    int C = (h->chain_cluster > 1 && S.n_mol >= 64) ? std::min(h->chain_cluster, CHAINC_MAXC) : 1;
    bool sliced = false;
    size_t smem = chain_smem_bytes(S.n_mol, h->US, recip ? S.nkvecs : 0);             // every CTA holds the whole system
    const bool whole_fits = smem + 16 * 1024 <= (size_t)max_optin && S.n_mol <= CHAIN_MAXIT * CHAIN_THREADS;
    if (!whole_fits) {                                                                  // a slice per CTA of an 8-SM cluster
        if (S.n_mol < 64) FAIL(MMC_EINVAL, "device loop: the system does not fit one SM's shared memory; use mmc_loop_run");
        C = CHAINC_MAXC; sliced = true;
        const int held = (S.n_mol + C - 1) / C + 1;
        smem = chain_smem_bytes(held, h->US, recip ? S.nkvecs : 0);
        if (smem + 16 * 1024 > (size_t)max_optin || held > CHAINS_MAXIT * CHAINS_WORKERS)
            FAIL(MMC_EINVAL, "device loop: the system does not fit the cluster's shared memory; use mmc_loop_run");
    } else if (C > 1 && (S.n_mol + C - 1) / C > CHAINS_MAXIT * CHAINS_WORKERS) C = 1;
    // a lone molecule: the block kernels prepare the next trial while the current one is evaluated, which presumes another
    // molecule to move next; the per-move kernels give the same record (and there is no pair work to batch)
    if (S.n_mol < 2) return mmc_loop_run(h, p, com, quat, db, uniforms, n_uniforms, n_moves, e0, v0, accepted, delta_out, st);
    if ((rc = flush_pending(h))) return rc;
    // one device block: [uniforms | quat (C replicas) | db | delta | out | accepted]
    const size_t n_q1 = 4 * (size_t)S.n_mol, n_q = n_q1 * C, n_db = 3 * (size_t)S.n_sites;
    const size_t off_q = (size_t)n_uniforms, off_db = off_q + n_q, off_delta = off_db + n_db, off_out = off_delta + (size_t)n_moves;
    const size_t out_doubles = (sizeof(ChainOut) + 7) / 8;
    const size_t bytes = (off_out + out_doubles) * sizeof(double) + (size_t)n_moves + 16;
    if (bytes > h->chain_bytes) {            // grown in powers of two from 4 MiB: a cudaFree + cudaMalloc between two blocks can cost
        size_t cap = 4u << 20;               // more than the block itself
        while (cap < bytes) cap <<= 1;
        dfree(h->d_chain);
        CK(cudaMalloc(&h->d_chain, cap));
        h->chain_bytes = cap;
    }
    double *d = reinterpret_cast<double *>(h->d_chain);
    unsigned char *d_acc = reinterpret_cast<unsigned char *>(d + off_out + out_doubles);
    CK(cudaMemcpyAsync(d, uniforms, sizeof(double) * (size_t)n_uniforms, cudaMemcpyHostToDevice, h->stream));
    for (int r = 0; r < C; ++r)
        CK(cudaMemcpyAsync(d + off_q + r * n_q1, quat, sizeof(double) * n_q1, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + off_db, db, sizeof(double) * n_db, cudaMemcpyHostToDevice, h->stream));
    ChainArgs A{};
    A.n_moves = n_moves; A.n_uniforms = n_uniforms;
    A.style_qq = p->style != MMC_STYLE_LJ_ONLY; A.style_recip = recip;
    A.adjust = p->adjust; A.cur = h->cur;
    A.temperature = p->temperature; A.inv_temperature = 1.0 / p->temperature; A.dr_max = p->dr_max; A.dphi_max = p->dphi_max;
    A.p_trans = p->p_trans; A.p_rot = p->p_rot; A.e0 = e0; A.v0 = v0;
    A.uniforms = d; A.quat = d + off_q; A.db = d + off_db; A.delta = d + off_delta;
    A.out = reinterpret_cast<ChainOut *>(d + off_out); A.accepted = d_acc;
#define MMC_CHAIN_LAUNCH(SS, DD)                                                                                     \
    if (C > 1) {                                                                                                     \
        auto kern = sliced ? k_chains<SS, DD, true> : k_chains<SS, DD, false>;                                        \
        CK(ensure_smem_optin((const void *)kern, (int)smem));                      \
        cudaLaunchConfig_t lc{};                                                                                     \
        lc.gridDim = dim3(C); lc.blockDim = dim3(CHAINC_THREADS); lc.dynamicSmemBytes = smem; lc.stream = h->stream; \
        cudaLaunchAttribute at[1];                                                                                   \
        at[0].id = cudaLaunchAttributeClusterDimension;                                                              \
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                          \
        lc.attrs = at; lc.numAttrs = 1;                                                                              \
        CK(cudaLaunchKernelEx(&lc, kern, h->S, A, h->move_poly));                                                    \
    } else {                                                                                                         \
        CK(ensure_smem_optin((const void *)k_chain<SS, DD>, (int)smem));           \
        k_chain<SS, DD><<<1, CHAIN_THREADS, smem, h->stream>>>(h->S, A, h->move_poly);                               \
    }
    const int deg = A.style_qq ? h->move_poly.deg : 0;
    if (h->US == 3) {          // water-like: erf polynomial degree at compile time
        switch (deg) {
            case 0: MMC_CHAIN_LAUNCH(3, 0) break;
            case 8: MMC_CHAIN_LAUNCH(3, 8) break;
            case 12: MMC_CHAIN_LAUNCH(3, 12) break;
            case 16: MMC_CHAIN_LAUNCH(3, 16) break;
            case 20: MMC_CHAIN_LAUNCH(3, 20) break;
            case 24: MMC_CHAIN_LAUNCH(3, 24) break;
            case 32: MMC_CHAIN_LAUNCH(3, 32) break;
            default: MMC_CHAIN_LAUNCH(3, -1) break;
        }
    } else if (deg == 0) {
        switch (h->US) {
            case 1: MMC_CHAIN_LAUNCH(1, 0) break;
            case 2: MMC_CHAIN_LAUNCH(2, 0) break;
            default: MMC_CHAIN_LAUNCH(4, 0) break;
        }
    } else {
        switch (h->US) {
            case 1: MMC_CHAIN_LAUNCH(1, -1) break;
            case 2: MMC_CHAIN_LAUNCH(2, -1) break;
            default: MMC_CHAIN_LAUNCH(4, -1) break;
        }
    }
#undef MMC_CHAIN_LAUNCH
    LAUNCH_CHECK();
    ChainOut o{};
    std::vector<double4> hc(S.n_mol);
    CK(cudaMemcpyAsync(&o, A.out, sizeof(o), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hc.data(), S.com, sizeof(double4) * S.n_mol, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(quat, A.quat, sizeof(double) * n_q1, cudaMemcpyDeviceToHost, h->stream));
    if (accepted && n_moves) CK(cudaMemcpyAsync(accepted, d_acc, (size_t)n_moves, cudaMemcpyDeviceToHost, h->stream));
    if (delta_out && n_moves) CK(cudaMemcpyAsync(delta_out, A.delta, sizeof(double) * (size_t)n_moves, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int m = 0; m < S.n_mol; ++m) { com[3 * m] = hc[m].x; com[3 * m + 1] = hc[m].y; com[3 * m + 2] = hc[m].z; }
    h->cur = o.cur;
    h->new_valid = false;
    h->state_version++;
    if (std::getenv("MMC_CHAIN_DEBUG") && o.n_moves > 0)
        std::fprintf(stderr, "k_chain cycles/move: step0 %.0f  gate %.0f  compact %.0f  pairs %.0f  B4wait %.0f  decide %.0f\n",
                     (double)o.phase_cycles[0] / o.n_moves, (double)o.phase_cycles[1] / o.n_moves, (double)o.phase_cycles[2] / o.n_moves,
                     (double)o.phase_cycles[3] / o.n_moves, (double)o.phase_cycles[4] / o.n_moves, (double)o.phase_cycles[5] / o.n_moves);
    h->cnt.trial_moves += o.n_moves; h->cnt.commits += o.n_accepted; h->cnt.overlap_events += o.n_overlap;
    std::memset(st, 0, sizeof(*st));
    st->n_moves = o.n_moves; st->n_accepted = o.n_accepted; st->n_overlap = o.n_overlap; st->uniforms_used = o.uniforms_used;
    st->trans_attempt = o.trans_attempt; st->trans_accept = o.trans_accept;
    st->rot_attempt = o.rot_attempt; st->rot_accept = o.rot_accept;
    st->dr_max = o.dr_max; st->dphi_max = o.dphi_max; st->total_energy = o.total_energy; st->total_virial = o.total_virial;
    return o.ret;
}

extern "C" int mmc_loop_run_atoms(mmc_handle *h, double temperature, double dr_max, double *r,
                                  const double *uniforms, int64_t n_uniforms, int64_t n_moves, double e0, double v0,
                                  uint8_t *accepted, double *delta_out, mmc_loop_stats *st)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
    if (!r || !uniforms || !st) FAIL(MMC_EINVAL, "null argument");
    UStream us{uniforms, n_uniforms, 0, false};
    std::memset(st, 0, sizeof(*st));
    st->total_energy = e0; st->total_virial = v0;
    const int n = h->At.n;
    const double box = h->At.box;
    int ret = 0, rc;
    for (int64_t m = 0; m < n_moves; ++m) {              // Monatomic/mainMonatomic.jl:373-413
        const int i = (int)(m % n);
        const int64_t pos_m = us.pos;
        double rnew[3];
        const double z0 = us.next(), z1 = us.next(), z2 = us.next();
        if (us.dry) { us.pos = pos_m; ret = 1; break; }  // not enough uniforms for this move: it never happened
        rnew[0] = r[3 * i] + (z0 - 0.5) * dr_max;
        rnew[1] = r[3 * i + 1] + (z1 - 0.5) * dr_max;
        rnew[2] = r[3 * i + 2] + (z2 - 0.5) * dr_max;
        pbc3(rnew, box);
        mmc_trial_result t;
        if ((rc = mmc_trial_atom(h, i + 1, rnew, &t))) return rc;
        const double delta = t.lj_new - t.lj_old;
        const bool acc = metropolis(delta / temperature, us);
        if (us.dry) {
            if ((rc = mmc_reject(h))) return rc;
            us.pos = pos_m; ret = 1; break;
        }
        if (acc) {
            st->total_energy += delta;
            st->total_virial += (t.lj_vir_new - t.lj_vir_old);
            st->n_accepted += 1;
            r[3 * i] = rnew[0]; r[3 * i + 1] = rnew[1]; r[3 * i + 2] = rnew[2];
            if ((rc = mmc_accept(h))) return rc;
        } else if ((rc = mmc_reject(h))) return rc;
        if (accepted) accepted[m] = acc ? 1 : 0;
        if (delta_out) delta_out[m] = delta;
        st->n_moves = m + 1;
        st->trans_attempt += 1; st->trans_accept += acc ? 1 : 0;
    }
    st->uniforms_used = us.pos;
    st->dr_max = dr_max;
    return ret;
}

// Monatomic/mainMonatomic.jl:373-413 for a block of moves in one launch on a thread-block cluster
// (kernels_chain.cuh k_chain_atoms).  Arguments, draw order, return codes and statistics of mmc_loop_run_atoms.
extern "C" int mmc_loop_run_atoms_device(mmc_handle *h, double temperature, double dr_max, double *r,
                                         const double *uniforms, int64_t n_uniforms, int64_t n_moves, double e0, double v0,
                                         uint8_t *accepted, double *delta_out, mmc_loop_stats *st)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
    if (!r || !uniforms || !st || n_moves < 0 || n_uniforms < 0) FAIL(MMC_EINVAL, "bad argument");
    if (h->trial_pending) FAIL(MMC_ESTATE, "a trial move is pending");
    const int n = h->At.n;
    int C = (n >= 4096 && h->chain_cluster_atoms > 8) ? CHAINA_MAXC : CHAINC_MAXC;      // 16 CTAs need the non-portable cluster size
    if (n < 64) FAIL(MMC_EINVAL, "device loop: at least 64 atoms");
    size_t smem = sizeof(float4) * (size_t)((n + CHAINC_MAXC - 1) / CHAINC_MAXC + 1);     // sized for the 8-CTA fallback
    int dev = 0, max_optin = 0, rc;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem + 8 * 1024 > (size_t)max_optin) FAIL(MMC_EINVAL, "device loop: the slice of atoms does not fit one SM's shared memory; use mmc_loop_run_atoms");
    if ((rc = flush_pending(h))) return rc;
    const size_t off_delta = (size_t)n_uniforms, off_out = off_delta + (size_t)n_moves;
    const size_t out_doubles = (sizeof(ChainOut) + 7) / 8;
    const size_t bytes = (off_out + out_doubles) * sizeof(double) + (size_t)n_moves + 16;
    if (bytes > h->chain_bytes) {            // grown in powers of two from 4 MiB: a cudaFree + cudaMalloc between two blocks can cost
        size_t cap = 4u << 20;               // more than the block itself
        while (cap < bytes) cap <<= 1;
        dfree(h->d_chain);
        CK(cudaMalloc(&h->d_chain, cap));
        h->chain_bytes = cap;
    }
    double *d = reinterpret_cast<double *>(h->d_chain);
    unsigned char *d_acc = reinterpret_cast<unsigned char *>(d + off_out + out_doubles);
    CK(cudaMemcpyAsync(d, uniforms, sizeof(double) * (size_t)n_uniforms, cudaMemcpyHostToDevice, h->stream));
    ChainAtomArgs A{};
    A.n_moves = n_moves; A.n_uniforms = n_uniforms;
    A.temperature = temperature; A.inv_temperature = 1.0 / temperature; A.dr_max = dr_max; A.e0 = e0; A.v0 = v0;
    {   // conservative FP32 gate: |d²_f32 − d²| <= 2·sqrt(3)·rc·δ + 3δ² with δ = 5·L·2^-24 (float positions <= L + dr,
        // their difference, the ±L image shift); 4x safety, and one ulp for the `!(r² > r_cut²)` equality
        const double L = h->At.box, rcut = h->At.rc, del = 5.0 * (L + dr_max) / 16777216.0;
        const double margin = 4.0 * (2.0 * 1.7320508075688772 * rcut * del + 3.0 * del * del + 3.6e-7 * rcut * rcut);
        A.gate_rc2f = std::nextafterf((float)(rcut * rcut + margin), INFINITY);
    }
    A.uniforms = d; A.delta = d + off_delta; A.out = reinterpret_cast<ChainOut *>(d + off_out); A.accepted = d_acc;
    CK(ensure_smem_optin((const void *)k_chain_atoms, (int)smem));
    static int nonportable_ok = -1;          // (set once: see ensure_smem_optin)
    if (C > 8 && nonportable_ok < 0)
        nonportable_ok = cudaFuncSetAttribute(k_chain_atoms, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : 0;
    if (C > 8 && !nonportable_ok) {
        cudaGetLastError();
        C = CHAINC_MAXC;
    }
    for (;;) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(C); lc.blockDim = dim3(CHAINA_THREADS); lc.dynamicSmemBytes = smem; lc.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&lc, k_chain_atoms, h->At, A);
        if (le == cudaSuccess) break;
        cudaGetLastError();
        if (C > 8) { C = CHAINC_MAXC; continue; }            // 16-CTA cluster not schedulable here: portable size
        h->err = std::string("k_chain_atoms launch: ") + cudaGetErrorString(le);
        return MMC_ECUDA;
    }
    LAUNCH_CHECK();
    ChainOut o{};
    std::vector<double4> hr(n);
    CK(cudaMemcpyAsync(&o, A.out, sizeof(o), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hr.data(), h->At.r, sizeof(double4) * n, cudaMemcpyDeviceToHost, h->stream));
    if (accepted && n_moves) CK(cudaMemcpyAsync(accepted, d_acc, (size_t)n_moves, cudaMemcpyDeviceToHost, h->stream));
    if (delta_out && n_moves) CK(cudaMemcpyAsync(delta_out, A.delta, sizeof(double) * (size_t)n_moves, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < n; ++k) { r[3 * k] = hr[k].x; r[3 * k + 1] = hr[k].y; r[3 * k + 2] = hr[k].z; }
    h->cnt.trial_moves += o.n_moves; h->cnt.commits += o.n_accepted;
    std::memset(st, 0, sizeof(*st));
    st->n_moves = o.n_moves; st->n_accepted = o.n_accepted; st->uniforms_used = o.uniforms_used;
    st->trans_attempt = o.trans_attempt; st->trans_accept = o.trans_accept;
    st->dr_max = o.dr_max; st->total_energy = o.total_energy; st->total_virial = o.total_virial;
    return o.ret;
}

extern "C" int mmc_host_register(void *ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return MMC_EINVAL;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return MMC_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return MMC_ECUDA; }
    return MMC_OK;
}

extern "C" int mmc_host_unregister(void *ptr)
{
    if (!ptr) return MMC_EINVAL;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorHostMemoryNotRegistered ? MMC_ESTATE : MMC_ECUDA; }
    return MMC_OK;
}

extern "C" int mmc_julia_rand(uint64_t seed, int64_t skip, double *out, int64_t n)
{
    if (skip < 0 || n < 0 || (n > 0 && !out)) return MMC_EINVAL;
    julia_mt::State s;
    julia_mt::seed(s, seed);
    for (int64_t k = 0; k < skip; ++k) (void)julia_mt::next(s);
    for (int64_t k = 0; k < n; ++k) out[k] = julia_mt::next(s);
    return MMC_OK;
}
