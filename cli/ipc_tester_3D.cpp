// ipc_tester_3D -c <config.yaml> — drop-in for /root/reference/examples/ipc_tester_3D.cpp:10-35. Parses SE(3) configs and
// graphs; the SE(3) CUDA path is not built in this revision, so the run stops with IPC_ERR_UNSUPPORTED after --parse-only work.
#include "ipc_host.hpp"
int main(int argc, char** argv) { return ipc_host::tester_main(argc, argv, 3); }
