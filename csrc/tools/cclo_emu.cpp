// cclo_emu: one collective engine as a stand-alone process.
//
//   cclo_emu --rank R --world W [--addr 127.0.0.1] [--base-port 5500] [--ctrl-port 6500+R]
//            [--mem-mb 256] [--no-kernel-loopback] [--log-level N]
//
// The engine talks to the other ranks' engines over loopback TCP (rank r listens on base-port + r) and
// serves exactly one driver (accl::emu::RemoteDevice) on the control port; it exits when the driver
// sends SHUTDOWN or disconnects.  Same role and flags as the reference's emulator executable
// (test/model/emulator/cclo_emu.cpp:510-537: -s world, -r rank, -p start port, -b kernel loopback,
// -l log level); `python -m accl_b200.models.emulator --engines` is the launcher (reference run.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "accl/common.hpp"
#include "accl/emu/remote.hpp"

using namespace accl;

int main(int argc, char **argv) {
  int rank = -1, world = -1, base_port = 5500, ctrl_port = -1, mem_mb = 256, log_level = -1;
  bool loopback = true;
  std::string addr = "127.0.0.1";
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char * {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "cclo_emu: %s needs a value\n", a.c_str());
        std::exit(2);
      }
      return argv[++i];
    };
    if (a == "--rank" || a == "-r") rank = std::atoi(next());
    else if (a == "--world" || a == "-s") world = std::atoi(next());
    else if (a == "--addr") addr = next();
    else if (a == "--base-port" || a == "-p") base_port = std::atoi(next());
    else if (a == "--ctrl-port") ctrl_port = std::atoi(next());
    else if (a == "--mem-mb") mem_mb = std::atoi(next());
    else if (a == "--no-kernel-loopback") loopback = false;
    else if (a == "--log-level" || a == "-l") log_level = std::atoi(next());
    else {
      std::fprintf(stderr, "usage: cclo_emu --rank R --world W [--addr A] [--base-port P] [--ctrl-port C] [--mem-mb M] "
                           "[--no-kernel-loopback] [--log-level N]\n");
      return a == "--help" || a == "-h" ? 0 : 2;
    }
  }
  if (rank < 0 || world <= 0 || rank >= world) {
    std::fprintf(stderr, "cclo_emu: --rank and --world are required\n");
    return 2;
  }
  if (ctrl_port < 0) ctrl_port = base_port + 1000 + rank;
  if (log_level >= 0) setenv("ACCL_LOG_LEVEL", std::to_string(log_level).c_str(), 1);
  try {
    auto fabric = std::make_shared<emu::SocketFabric>(rank, world, addr, base_port);
    auto engine = std::make_shared<emu::Engine>(rank, world, fabric, static_cast<size_t>(mem_mb) << 20,
                                                static_cast<size_t>(mem_mb) << 20);
    engine->set_kernel_loopback(loopback);
    emu::EngineServer server(engine, addr, ctrl_port);
    std::fprintf(stderr, "cclo_emu rank %d/%d: fabric %s:%d control %s:%d\n", rank, world, addr.c_str(), base_port + rank,
                 addr.c_str(), ctrl_port);
    server.serve();
  } catch (const std::exception &e) {
    std::fprintf(stderr, "cclo_emu rank %d: %s\n", rank, e.what());
    return 1;
  }
  return 0;
}
