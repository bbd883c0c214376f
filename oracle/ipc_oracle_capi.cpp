// ipc_oracle_capi.cpp — flat C API over the CPU ORACLE (test infrastructure; see ipc_oracle.hpp).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
#include "ipc_oracle.hpp"

#include <atomic>
#include <thread>

using namespace ipc_oracle;

extern "C" {

struct orc_report {            // mirrors CheckReport
    int accepted, slow_path, lo, hi, n_eset, n_cluster, iterations, evals, result, _pad;
    double max_chi2, cand_chi2, sum_chi2;
};

struct orc_config {
    double s_factor, fast_reject_th, slow_reject_th;
    int fast_reject_iter_base, slow_reject_iter_base;
};
}

namespace {

template <class G> Edge<G> make_edge(int from, int to, const double* meas, const double* info) {
    Edge<G> e; e.from = from; e.to = to; e.set_meas(G::from_flat(meas));
    for (int q = 0; q < G::D * G::D; ++q) e.info.a[q] = info[q];
    return e;
}
void fill(orc_report* o, const CheckReport& r) {
    if (!o) return;
    o->accepted = r.accepted; o->slow_path = r.slow_path; o->lo = r.lo; o->hi = r.hi; o->n_eset = r.n_eset;
    o->n_cluster = r.n_cluster; o->iterations = r.iterations; o->evals = r.evals; o->result = r.result; o->_pad = 0;
    o->max_chi2 = r.max_chi2; o->cand_chi2 = r.cand_chi2; o->sum_chi2 = r.sum_chi2;
}

struct HandleBase { int dim; virtual ~HandleBase() {} };
template <class G> struct Handle : HandleBase {
    IPC<G> ipc;
    std::vector<typename G::Pose> pristine;
    Handle(int n, std::vector<Edge<G>> od, const IpcConfig& c) : ipc(n, std::move(od), c) { dim = G::D == 3 ? 2 : 3; pristine = ipc.est; }
};

template <class G> HandleBase* create(int n_poses, const double* odom_meas, const double* odom_info, const orc_config* c) {
    std::vector<Edge<G>> od;
    for (int j = 0; j + 1 < n_poses; ++j) od.push_back(make_edge<G>(j, j + 1, odom_meas + j * G::MEAS, odom_info + j * G::D * G::D));
    IpcConfig ic; ic.s_factor = c->s_factor; ic.fast_reject_th = c->fast_reject_th; ic.slow_reject_th = c->slow_reject_th;
    ic.fast_reject_iter_base = c->fast_reject_iter_base; ic.slow_reject_iter_base = c->slow_reject_iter_base;
    return new Handle<G>(n_poses, std::move(od), ic);
}

template <class G> void batch(Handle<G>* h, const int* lfrom, const int* lto, const double* lmeas, const double* linfo,
                              int n_checks, const int* cptr, const int* cidx, int n_threads, unsigned char* out_accept, orc_report* out_rep) {
    std::atomic<int> next{0};
    auto work = [&]() {
        IPC<G> local = h->ipc;           // private copy (odom already scaled)
        local.cns.clear();
        for (;;) {
            int c = next.fetch_add(1);
            if (c >= n_checks) break;
            local.est = h->pristine;
            local.cns.clear();
            int b = cptr[c], e = cptr[c + 1];
            for (int q = b; q + 1 < e; ++q) { int l = cidx[q]; local.addEdgeToCnS(make_edge<G>(lfrom[l], lto[l], lmeas + l * G::MEAS, linfo + l * G::D * G::D)); }
            int l = cidx[e - 1];
            CheckReport rep;
            bool ok = local.agreementCheck(make_edge<G>(lfrom[l], lto[l], lmeas + l * G::MEAS, linfo + l * G::D * G::D), &rep);
            out_accept[c] = ok ? 1 : 0;
            if (out_rep) fill(out_rep + c, rep);
        }
    };
    if (n_threads <= 1) { work(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

void* orc_create(int dim, int n_poses, const double* odom_meas, const double* odom_info, const orc_config* cfg) {
    if (dim == 2) return create<G2>(n_poses, odom_meas, odom_info, cfg);
    if (dim == 3) return create<G3>(n_poses, odom_meas, odom_info, cfg);
    return nullptr;
}
void orc_destroy(void* h) { delete static_cast<HandleBase*>(h); }

void orc_set_noise_exit(void* hv, double eps) {   // eps > 0 enables the retry shortcut with that threshold, 0 disables it
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) { auto& i = static_cast<Handle<G2>*>(hb)->ipc; i.noise_exit = eps > 0; i.noise_eps = eps; }
    else { auto& i = static_cast<Handle<G3>*>(hb)->ipc; i.noise_exit = eps > 0; i.noise_eps = eps; }
}

int orc_agreement_check(void* hv, int from, int to, const double* meas, const double* info, orc_report* rep) {
    auto* hb = static_cast<HandleBase*>(hv);
    CheckReport r; bool ok;
    if (hb->dim == 2) ok = static_cast<Handle<G2>*>(hb)->ipc.agreementCheck(make_edge<G2>(from, to, meas, info), &r);
    else ok = static_cast<Handle<G3>*>(hb)->ipc.agreementCheck(make_edge<G3>(from, to, meas, info), &r);
    fill(rep, r);
    return ok ? 1 : 0;
}
void orc_add_edge(void* hv, int from, int to, const double* meas, const double* info) {
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) static_cast<Handle<G2>*>(hb)->ipc.addEdgeToCnS(make_edge<G2>(from, to, meas, info));
    else static_cast<Handle<G3>*>(hb)->ipc.addEdgeToCnS(make_edge<G3>(from, to, meas, info));
}
int orc_remove_edge(void* hv, int from, int to) {
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) return static_cast<Handle<G2>*>(hb)->ipc.removeEdgeFromCnS(from, to);
    return static_cast<Handle<G3>*>(hb)->ipc.removeEdgeFromCnS(from, to);
}
int orc_consensus_size(void* hv) {
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) return (int)static_cast<Handle<G2>*>(hb)->ipc.cns.size();
    return (int)static_cast<Handle<G3>*>(hb)->ipc.cns.size();
}
void orc_get_consensus(void* hv, int* from_to) {   // 2 ints per edge
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) { auto& c = static_cast<Handle<G2>*>(hb)->ipc.cns; for (size_t i = 0; i < c.size(); ++i) { from_to[2 * i] = c[i].from; from_to[2 * i + 1] = c[i].to; } }
    else { auto& c = static_cast<Handle<G3>*>(hb)->ipc.cns; for (size_t i = 0; i < c.size(); ++i) { from_to[2 * i] = c[i].from; from_to[2 * i + 1] = c[i].to; } }
}
// the final full-graph optimisation (src/simulation.cpp:50-65); returns chi2, *iterations = Dogleg iterations executed
double orc_final_optimize(void* hv, int max_iterations, int* iterations) {
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) return static_cast<Handle<G2>*>(hb)->ipc.finalOptimize(max_iterations, iterations);
    return static_cast<Handle<G3>*>(hb)->ipc.finalOptimize(max_iterations, iterations);
}
void orc_get_poses(void* hv, double* out) {        // 3 (x y th) or 7 (x y z qx qy qz qw) per pose
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) { auto& e = static_cast<Handle<G2>*>(hb)->ipc.est; for (size_t i = 0; i < e.size(); ++i) G2::to_flat(e[i], out + 3 * i); }
    else { auto& e = static_cast<Handle<G3>*>(hb)->ipc.est; for (size_t i = 0; i < e.size(); ++i) G3::to_flat(e[i], out + 7 * i); }
}
// Independent checks from the dead-reckoned state. Check c lists loop indices cidx[cptr[c]..cptr[c+1]):
// all but the last are put in a fresh consensus set (addEdgeToCnS), the last is agreementCheck'ed.
void orc_check_batch(void* hv, const int* lfrom, const int* lto, const double* lmeas, const double* linfo,
                     int n_checks, const int* cptr, const int* cidx, int n_threads, unsigned char* out_accept, orc_report* out_rep) {
    auto* hb = static_cast<HandleBase*>(hv);
    if (hb->dim == 2) batch<G2>(static_cast<Handle<G2>*>(hb), lfrom, lto, lmeas, linfo, n_checks, cptr, cidx, n_threads, out_accept, out_rep);
    else batch<G3>(static_cast<Handle<G3>*>(hb), lfrom, lto, lmeas, linfo, n_checks, cptr, cidx, n_threads, out_accept, out_rep);
}

// unit-test hooks: error and Jacobians of one edge, oplus
void orc_edge_eval(int dim, const double* meas, const double* xi, const double* xj, double* err, double* Ji, double* Jj) {
    if (dim == 2) {
        auto z = G2::from_flat(meas); auto a = G2::from_flat(xi), b = G2::from_flat(xj);
        auto e = G2::error(z.inverse(), a, b); Mat<3> A, B; G2::jacobians(z.inverse(), a, b, A, B);
        for (int i = 0; i < 3; ++i) err[i] = e[i];
        for (int i = 0; i < 9; ++i) { Ji[i] = A.a[i]; Jj[i] = B.a[i]; }
    } else {
        auto z = G3::from_flat(meas); auto a = G3::from_flat(xi), b = G3::from_flat(xj);
        auto e = G3::error(z.inverse(), a, b); Mat<6> A, B; G3::jacobians(z.inverse(), a, b, A, B);
        for (int i = 0; i < 6; ++i) err[i] = e[i];
        for (int i = 0; i < 36; ++i) { Ji[i] = A.a[i]; Jj[i] = B.a[i]; }
    }
}
void orc_oplus(int dim, const double* x, const double* u, double* out) {
    if (dim == 2) { auto p = G2::from_flat(x); G2::oplus(p, u); G2::to_flat(p, out); }
    else { auto p = G3::from_flat(x); G3::oplus(p, u); G3::to_flat(p, out); }
}
void orc_compose(int dim, const double* a, const double* b, double* out) {
    if (dim == 2) G2::to_flat(G2::compose(G2::from_flat(a), G2::from_flat(b)), out);
    else G3::to_flat(G3::compose(G3::from_flat(a), G3::from_flat(b)), out);
}
void orc_inverse(int dim, const double* a, double* out) {
    if (dim == 2) G2::to_flat(G2::inverse(G2::from_flat(a)), out);
    else G3::to_flat(G3::inverse(G3::from_flat(a)), out);
}
}
