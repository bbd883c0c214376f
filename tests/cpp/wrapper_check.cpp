// CPU-side check of include/ipc_b200.hpp (tests/test_cpp_wrapper.py): every member of both instantiations compiles against the C ABI,
// argument errors surface as ipc_b200::Error before any device work, and without a CUDA device construction fails loudly
// (IPC_ERR_CUDA) — there is no CPU path behind the class.
#include <cstdio>
#include <cstring>

#include "ipc_b200.hpp"

template class ipc_b200::IPC<2>;
template class ipc_b200::IPC<3>;

static ipc_b200::Edge odom2(int j) {
    ipc_b200::Edge e;
    e.from = j; e.to = j + 1;
    e.meas = {1.0, 0.0, 0.1};
    e.info = {10, 0, 0, 0, 10, 0, 0, 0, 5};
    return e;
}

int main() {
    using namespace ipc_b200;
    Config cfg;
    cfg.s_factor = 10; cfg.fast_reject_th = 6.251; cfg.slow_reject_th = 11.345; cfg.fast_reject_iter_base = 50; cfg.slow_reject_iter_base = 100;
    int fails = 0;
    {   // a gap in the odometry chain is an argument error raised by the wrapper itself
        std::vector<Edge> od = {odom2(0), odom2(2)};
        try { IPC2D ipc(od, cfg); std::printf("FAIL: gap accepted\n"); ++fails; }
        catch (const Error& e) { if (e.code != IPC_ERR_ARG) { std::printf("FAIL: gap -> code %d\n", e.code); ++fails; } }
    }
    {   // wrong measurement size
        std::vector<Edge> od = {odom2(0), odom2(1)};
        od[1].meas.push_back(0.0);
        try { IPC2D ipc(od, cfg); std::printf("FAIL: bad size accepted\n"); ++fails; }
        catch (const Error& e) { if (e.code != IPC_ERR_ARG) { std::printf("FAIL: bad size -> code %d\n", e.code); ++fails; } }
    }
    {   // a well-formed graph: on a box without a GPU the constructor must throw IPC_ERR_CUDA (no CPU fallback); on a GPU box it works
        std::vector<Edge> od;
        for (int j = 0; j < 8; ++j) od.push_back(odom2(j));
        const int ndev = ipc_device_count();
        try {
            IPC2D ipc(od, cfg);
            if (ndev <= 0) { std::printf("FAIL: constructed without a device\n"); ++fails; }
            else {
                Edge loop; loop.from = 0; loop.to = 7; loop.meas = {7.0 * 0.9, 2.0, 0.7}; loop.info = {10, 0, 0, 0, 10, 0, 0, 0, 5};
                const bool ok = ipc.agreementCheck(loop);
                std::printf("device run: agreementCheck -> %d, consensus %zu\n", (int)ok, ipc.getMaxConsensusSet().size());
            }
        } catch (const Error& e) {
            if (ndev > 0 || e.code != IPC_ERR_CUDA) { std::printf("FAIL: create -> code %d (%s)\n", e.code, e.what()); ++fails; }
            else std::printf("no device: %s\n", e.what());
        }
    }
    std::printf(fails ? "WRAPPER_FAIL\n" : "WRAPPER_OK\n");
    return fails ? 1 : 0;
}
