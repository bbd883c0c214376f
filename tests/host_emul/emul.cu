// Host emulation of the CUDA check (TEST INFRASTRUCTURE): compiles ipc_b200/csrc/chain_se2.cuh with NT = 1 for the
// CPU so the arithmetic and the Dogleg control flow of the kernel can be validated against the oracle on a box without
// a GPU; emul_set_cta_threads(8 | 16) runs every check on that many cooperating OS threads instead (the kernel's block
// decomposition: segments, scratch slots, boundary vertices, block collectives). Never part of the product library.
#include <cstdio>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../ipc_b200/csrc/chain_se2.cuh"
#include "../../ipc_b200/csrc/host_state.hpp"

using namespace ipcb;

static int* g_diag = nullptr;
static int g_sd_fuse = 2;
static int g_cta = 1;
static int g_gst = 0;      // 1: run the checks with the global-memory state layout (tiles by step, StateAt<NT, true>)
extern "C" void emul_set_global_state(int v) { g_gst = v; }
extern "C" int emul_set_cta_threads(int nt) { if (nt != 1 && nt != 8 && nt != 16) return -1; g_cta = nt; return 0; }

struct SpinBarrier {
    std::atomic<int> count{0}, gen{0};
    int n = 1;
    void wait() {
        const int g = gen.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { count.store(0, std::memory_order_relaxed); gen.fetch_add(1, std::memory_order_acq_rel); }
        else while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
    }
    static void sync(void* p) { static_cast<SpinBarrier*>(p)->wait(); }
};

// one emulated CTA of NT threads working through the check list
template <int NT, bool UNI, class Out>
static void cta_group(int n_poses, const double* odom, const double* Du, const double* Vu, const LoopRec2* recs, int n_checks, const int* member,
                      const int* cand, const CheckParams& prm, bool want_info, std::atomic<int>& next, Out&& out) {
    const int capv = n_poses + 2, capg = scratch_slots<NT>(capv);
    const size_t st_doubles = std::max((size_t)CHAIN_STATE_ARRAYS * capv, (size_t)global_state_doubles(capv, NT));
    std::vector<double> buf(2 * st_doubles + (size_t)CHAIN_SCRATCH_ARRAYS * capg + CHAIN_SMALL_DOUBLES, 0.0);
    ChainMem M; double* p = buf.data();
    M.small = p; M.st = p + CHAIN_SMALL_DOUBLES; M.scr = M.st + 2 * st_doubles; M.capv = capv; M.capg = capg;
    M.stw = g_gst ? M.st + st_doubles : M.st; M.ring = nullptr;
    SpinBarrier bar; bar.n = NT;
    const bool gst = g_gst != 0;
    int cur = 0;
    auto body = [&](int tid) {
        HostCta cta{tid, &SpinBarrier::sync, &bar};
        host_cta() = &cta;
        for (;;) {
            if (tid == 0) cur = next.fetch_add(1);
            bar.wait();
            const int c = cur;
            if (c >= n_checks) break;
            CheckResult r;
            if (gst) run_check<NT, UNI, false, true>(M, odom, Du, Vu, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info, r);
            else run_check<NT, UNI>(M, odom, Du, Vu, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info, r);
            if (tid == 0) out(c, r);
            bar.wait();
        }
        host_cta() = nullptr;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < NT; ++t) th.emplace_back(body, t);
    body(0);
    for (auto& t : th) t.join();
}
extern "C" void emul_set_sd_fuse(int v) { g_sd_fuse = v; }
extern "C" void emul_set_diag(int* p) { g_diag = p; }
template <class PT> static int emul_impl(int n_poses, const double* odom_meas, const double* odom_info, double s_factor, int n_loops, const int* lfrom,
                                const int* lto, const double* lmeas, const double* linfo, int n_checks, const int* member, const int* cand,
                                double fast_th, double slow_th, int fast_iter, int slow_iter, double noise_eps, int speculate, int early_accept,
                                int want_info, int use_uni, int n_threads, unsigned char* verdict, ipc_check_info* info, int* sweeps) {
    HostState hs; std::string err;
    if (!hs.init(2, n_poses, odom_meas, odom_info, s_factor, err)) return -1;
    const int n_pad = (n_poses + 3) & ~1;
    const bool uni = hs.uniform_iso && use_uni;
    std::vector<double> soa; hs.build_odom_aos(uni, n_pad, soa);
    std::vector<LoopRec2> recs(n_loops);
    for (int i = 0; i < n_loops; ++i) { recs[i].from = lfrom[i]; recs[i].to = lto[i]; HostState::se2_edge_record(lmeas + 3 * i, linfo + 9 * i, 1.0, recs[i].meas, recs[i].D); HostState::inv_sym3_host(recs[i].D, recs[i].V); }
    CheckParams prm{fast_th, slow_th, fast_iter, slow_iter, noise_eps, 100, speculate, early_accept, g_sd_fuse};
    std::atomic<int> next{0};
    auto emit = [&](int c, const CheckResult& r) {
        verdict[c] = (unsigned char)r.verdict;
        if (info) { info[c].max_chi2 = r.max_chi2; info[c].cand_chi2 = r.cand_chi2; info[c].sum_chi2 = r.sum_chi2; info[c].iterations = r.iterations;
                    info[c].evals = r.evals; info[c].window_len = r.window_len; info[c].n_loops = r.n_loops; }
        if (sweeps) sweeps[c] = r.n_sweeps;
        if (g_diag) { g_diag[4 * c] = r.n_norm; g_diag[4 * c + 1] = r.n_sd; g_diag[4 * c + 2] = r.n_relin; g_diag[4 * c + 3] = r.n_blend; }
    };
    if (g_cta > 1) {
        const int groups = std::max(1, (n_threads < 1 ? 1 : n_threads) / g_cta);
        auto group = [&]() {
            const LoopRec2* rp = recs.data();
            if (g_cta == 8) { if (uni) cta_group<8, true>(n_poses, soa.data(), hs.Du, hs.Vu, rp, n_checks, member, cand, prm, want_info != 0, next, emit);
                              else cta_group<8, false>(n_poses, soa.data(), hs.Du, hs.Vu, rp, n_checks, member, cand, prm, want_info != 0, next, emit); }
            else { if (uni) cta_group<16, true>(n_poses, soa.data(), hs.Du, hs.Vu, rp, n_checks, member, cand, prm, want_info != 0, next, emit);
                   else cta_group<16, false>(n_poses, soa.data(), hs.Du, hs.Vu, rp, n_checks, member, cand, prm, want_info != 0, next, emit); }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < groups; ++t) th.emplace_back(group);
        for (auto& t : th) t.join();
        return 0;
    }
    auto work = [&]() {
        const int capv = n_poses + 2;
        const int capg = scratch_slots<1>(capv);
        const size_t st_doubles = std::max((size_t)CHAIN_STATE_ARRAYS * capv, (size_t)global_state_doubles(capv, 1));
        std::vector<double> buf(2 * st_doubles + (size_t)CHAIN_SCRATCH_ARRAYS * capg + CHAIN_SMALL_DOUBLES, 0.0);
        ChainMem M; double* p = buf.data();
        M.small = p; M.st = p + CHAIN_SMALL_DOUBLES; M.scr = M.st + 2 * st_doubles; M.capv = capv; M.capg = capg;
        M.stw = g_gst ? M.st + st_doubles : M.st; M.ring = nullptr;
        const bool gst = g_gst != 0;
        for (;;) {
            int c = next.fetch_add(1);
            if (c >= n_checks) break;
            CheckResult r;
            if (gst) { if (uni) run_check<1, true, false, true>(M, soa.data(), hs.Du, hs.Vu, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info != 0, r);
                       else run_check<1, false, false, true>(M, soa.data(), hs.Du, hs.Vu, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info != 0, r); }
            else if (uni) run_check<1, true>(M, soa.data(), hs.Du, hs.Vu, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info != 0, r);
            else run_check<1, false>(M, soa.data(), hs.Du, hs.Vu, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info != 0, r);
            emit(c, r);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
    return 0;
}

extern "C" int emul_check_batch(int n_poses, const double* odom_meas, const double* odom_info, double s_factor, int n_loops, const int* lfrom,
                                const int* lto, const double* lmeas, const double* linfo, int n_checks, const int* member, const int* cand,
                                double fast_th, double slow_th, int fast_iter, int slow_iter, double noise_eps, int speculate, int early_accept,
                                int want_info, int use_uni, int n_threads, unsigned char* verdict, ipc_check_info* info, int* sweeps, int prefix_f32) {
    if (prefix_f32) return emul_impl<float>(n_poses, odom_meas, odom_info, s_factor, n_loops, lfrom, lto, lmeas, linfo, n_checks, member, cand, fast_th, slow_th,
                                            fast_iter, slow_iter, noise_eps, speculate, early_accept, want_info, use_uni, n_threads, verdict, info, sweeps);
    return emul_impl<double>(n_poses, odom_meas, odom_info, s_factor, n_loops, lfrom, lto, lmeas, linfo, n_checks, member, cand, fast_th, slow_th,
                             fast_iter, slow_iter, noise_eps, speculate, early_accept, want_info, use_uni, n_threads, verdict, info, sweeps);
}

// ---- SE(3) ------------------------------------------------------------------------------------------------------------
#include "../../ipc_b200/csrc/chain_se3.cuh"
template <int NT, class Out>
static void cta_group3(int n_poses, const double* odom, const ipcb::se3::LoopRec3* recs, int n_checks, const int* member, const int* cand,
                       const CheckParams& prm, bool want_info, std::atomic<int>& next, Out&& out) {
    using namespace ipcb::se3;
    const int capv = std::max(n_poses + 2, NT);
    std::vector<double> buf(2 * (size_t)global_state3_doubles(capv, NT) + (size_t)CHAIN3_SCRATCH * capv + CHAIN3_SMALL_DOUBLES, 0.0);
    ChainMem3 M; double* p = buf.data();
    M.small = p; M.st = p + CHAIN3_SMALL_DOUBLES; M.scr = M.st + 2 * (size_t)global_state3_doubles(capv, NT); M.capv = capv;
    M.gst = g_gst; M.nt = NT; M.S = 1; M.cs = g_gst ? NT : 1; M.stw = g_gst ? M.st + global_state3_doubles(capv, NT) : M.st;
    SpinBarrier bar; bar.n = NT;
    int cur = 0;
    auto body = [&](int tid) {
        HostCta cta{tid, &SpinBarrier::sync, &bar};
        host_cta() = &cta;
        for (;;) {
            if (tid == 0) cur = next.fetch_add(1);
            bar.wait();
            const int c = cur;
            if (c >= n_checks) break;
            CheckResult r;
            run_check3<NT>(M, odom, &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info, r);
            if (tid == 0) out(c, r);
            bar.wait();
        }
        host_cta() = nullptr;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < NT; ++t) th.emplace_back(body, t);
    body(0);
    for (auto& t : th) t.join();
}
extern "C" int emul_check_batch3(int n_poses, const double* odom_meas, const double* odom_info, double s_factor, int n_loops, const int* lfrom,
                                 const int* lto, const double* lmeas, const double* linfo, int n_checks, const int* member, const int* cand,
                                 double fast_th, double slow_th, int fast_iter, int slow_iter, double noise_eps, int speculate, int early_accept,
                                 int want_info, int n_threads, unsigned char* verdict, ipc_check_info* info, int* sweeps) {
    using namespace ipcb::se3;
    HostState hs; std::string err;
    if (!hs.init(3, n_poses, odom_meas, odom_info, s_factor, err)) return -1;
    const int n_pad = (n_poses + 3) & ~1;
    std::vector<double> rec;
    if (!hs.build_odom_aos3(n_pad, rec)) return -2;
    std::vector<LoopRec3> recs(n_loops);
    for (int i = 0; i < n_loops; ++i) {
        double r[49];
        if (!HostState::se3_edge_record(lmeas + 7 * i, linfo + 36 * i, 1.0, r)) return -3;
        recs[i].from = lfrom[i]; recs[i].to = lto[i];
        for (int q = 0; q < 7; ++q) recs[i].zinv[q] = r[q];
        for (int q = 0; q < 21; ++q) { recs[i].Om[q] = r[7 + q]; recs[i].V[q] = r[28 + q]; }
    }
    CheckParams prm{fast_th, slow_th, fast_iter, slow_iter, noise_eps, 100, speculate, early_accept, g_sd_fuse};
    std::atomic<int> next{0};
    auto emit = [&](int c, const CheckResult& r) {
        verdict[c] = (unsigned char)r.verdict;
        if (info) { info[c].max_chi2 = r.max_chi2; info[c].cand_chi2 = r.cand_chi2; info[c].sum_chi2 = r.sum_chi2; info[c].iterations = r.iterations;
                    info[c].evals = r.evals; info[c].window_len = r.window_len; info[c].n_loops = r.n_loops; }
        if (sweeps) sweeps[c] = r.n_sweeps;
    };
    if (g_cta > 1) {
        auto group = [&]() { if (g_cta == 8) cta_group3<8>(n_poses, rec.data(), recs.data(), n_checks, member, cand, prm, want_info != 0, next, emit);
                             else cta_group3<16>(n_poses, rec.data(), recs.data(), n_checks, member, cand, prm, want_info != 0, next, emit); };
        const int groups = std::max(1, (n_threads < 1 ? 1 : n_threads) / g_cta);
        std::vector<std::thread> th;
        for (int t = 0; t < groups; ++t) th.emplace_back(group);
        for (auto& t : th) t.join();
        return 0;
    }
    auto work = [&]() {
        const int capv = n_poses + 2;
        std::vector<double> buf((size_t)(CHAIN3_STATE + CHAIN3_SCRATCH) * capv + CHAIN3_SMALL_DOUBLES, 0.0);
        ChainMem3 M; double* p = buf.data();
        M.small = p; M.st = p + CHAIN3_SMALL_DOUBLES; M.scr = M.st + (size_t)CHAIN3_STATE * capv; M.capv = capv;
        M.gst = 0; M.nt = 1; M.S = 1; M.cs = 1; M.stw = M.st;
        for (;;) {
            int c = next.fetch_add(1);
            if (c >= n_checks) break;
            CheckResult r;
            run_check3<1>(M, rec.data(), &recs[cand[c]], member[c] >= 0 ? &recs[member[c]] : nullptr, prm, want_info != 0, r);
            verdict[c] = (unsigned char)r.verdict;
            if (info) { info[c].max_chi2 = r.max_chi2; info[c].cand_chi2 = r.cand_chi2; info[c].sum_chi2 = r.sum_chi2; info[c].iterations = r.iterations;
                        info[c].evals = r.evals; info[c].window_len = r.window_len; info[c].n_loops = r.n_loops; }
            if (sweeps) sweeps[c] = r.n_sweeps;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
    return 0;
}
