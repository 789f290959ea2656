// context.cpp -- ipcl::initializeContext / terminateContext on the CUDA
// back-end (reference: ipcl/utils/context.cpp:40-86, where these bring the QAT
// devices up and down).
#include "ipcl/utils/context.hpp"

#include <algorithm>
#include <cctype>
#include <cstdlib>

#include "ipcl/utils/util.hpp"
#include "ipcl_b200.h"

namespace ipcl {

static bool g_gpu_up = false;

bool initializeContext(const std::string runtime_choice) {
  std::string c = runtime_choice;
  std::transform(c.begin(), c.end(), c.begin(),
                 [](unsigned char ch) { return std::toupper(ch); });
  ERROR_CHECK(c == "DEFAULT" || c == "CPU" || c == "QAT" || c == "HYBRID" ||
                  c == "GPU" || c == "B200",
              "initializeContext: unknown runtime choice " + runtime_choice);
  // IPCLB200_DEVICES=n (or "all"): one process drives n GPUs, every batch is
  // split over them (the slot of the reference's CPU/QAT split, mod_exp.cpp:702-731)
  const char* nd = std::getenv("IPCLB200_DEVICES");
  if (nd && *nd) {
    const int n = (std::string(nd) == "all") ? 0 : std::atoi(nd);
    DEVICE_CHECK(ipclb200_init_devices(n));
  } else {
    DEVICE_CHECK(ipclb200_init(-1));
  }
  g_gpu_up = true;
  return true;
}

bool terminateContext() {
  ipclb200_shutdown();
  g_gpu_up = false;
  return true;
}

bool isQATRunning() { return false; }
bool isQATActive() { return false; }
bool isGPURunning() { return g_gpu_up; }

}  // namespace ipcl
