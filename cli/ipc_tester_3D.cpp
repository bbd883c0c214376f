// ipc_tester_3D -c <config.yaml> — drop-in for /root/reference/examples/ipc_tester_3D.cpp:10-35 over libipc_b200.so
// (EDGE_SE3:QUAT / VERTEX_SE3:QUAT graphs, IPC<EdgeSE3, VertexSE3>).
#include "ipc_host.hpp"
int main(int argc, char** argv) { return ipc_host::tester_main(argc, argv, 3); }
