// ipc_tester_2D -c <config.yaml> — drop-in for /root/reference/examples/ipc_tester_2D.cpp:10-35 over libipc_b200.so.
#include "ipc_host.hpp"
int main(int argc, char** argv) { return ipc_host::tester_main(argc, argv, 2); }
