// Load time of a g2o file (reference: optimizer.load in src/utils.cpp:95-126 -> g2o's line-by-line istream reader):
// cli/ipc_host.hpp::loadG2O (whole file in one read, chunks cut at line boundaries, strtod tokeniser on worker threads) with 1 .. T
// threads, next to a plain getline + istringstream loader of the same records (the shape of the reference's reader) on the same file.
// Build / run:  g++ -O2 -std=c++17 -pthread -Icli -Iinclude scripts/bench_g2o_load.cpp -o /tmp/bench_g2o_load && /tmp/bench_g2o_load file.g2o 2
#include <chrono>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <thread>

#include "ipc_host.hpp"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static size_t load_getline(const std::string& path, int dim, size_t& n_edges) {
    std::ifstream f(path);
    std::string line, tag;
    size_t n_vertices = 0; n_edges = 0;
    const int nm = dim == 2 ? 3 : 7, ni = dim == 2 ? 6 : 21;
    std::vector<double> keep;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        if (!(is >> tag)) continue;
        if (tag.rfind("VERTEX", 0) == 0) { int id; is >> id; double x; for (int q = 0; q < nm; ++q) { is >> x; keep.push_back(x); } ++n_vertices; }
        else if (tag.rfind("EDGE", 0) == 0) { int a, b; is >> a >> b; double x; for (int q = 0; q < nm + ni; ++q) { is >> x; keep.push_back(x); } ++n_edges; }
    }
    return n_vertices + (keep.empty() ? 1 : 0);
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s file.g2o dim\n", argv[0]); return 2; }
    const std::string path = argv[1];
    const int dim = std::atoi(argv[2]);
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    const double mb = (double)f.tellg() / 1e6;
    const int hw = (int)std::thread::hardware_concurrency();
    std::printf("{\"file_MB\": %.2f, \"host_threads\": %d", mb, hw);
    size_t nv = 0, ne = 0;
    for (int nt : {1, 2, 4, 8, 16}) {
        if (nt > std::max(1, hw)) break;
        double best = 1e30;
        for (int rep = 0; rep < 5; ++rep) {
            const double t = now();
            ipc_host::Graph g = ipc_host::loadG2O(path, dim, nt);
            best = std::min(best, now() - t);
            nv = g.vertex_ids.size(); ne = g.edges.size();
        }
        std::printf(", \"loadG2O_%dthr_ms\": %.2f", nt, best * 1e3);
    }
    double best = 1e30; size_t ne2 = 0, nv2 = 0;
    for (int rep = 0; rep < 3; ++rep) { const double t = now(); nv2 = load_getline(path, dim, ne2); best = std::min(best, now() - t); }
    std::printf(", \"getline_istringstream_ms\": %.2f, \"vertices\": %zu, \"edges\": %zu, \"same_counts\": %s}\n", best * 1e3, nv, ne, (nv == nv2 && ne == ne2) ? "true" : "false");
    return 0;
}
