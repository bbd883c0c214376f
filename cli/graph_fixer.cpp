// graph_fixer — examples/graph_fixer.cpp of the reference: re-orient loop edges so that from < to. Host-only (no GPU work).
#include "ipc_host.hpp"
int main(int argc, char** argv) { return ipc_host::graph_fixer_main(argc, argv); }
