#include "ipcl/utils/util.hpp"

#include "ipcl_b200.h"

namespace ipcl {

void check_device_status(int rc, const char* what, const char* file, int line) {
  if (rc == IPCLB200_OK) return;
  throw std::runtime_error(build_log(
      file, line,
      std::string(what) + " failed with status " + std::to_string(rc) + ": " +
          ipclb200_last_error()));
}

}  // namespace ipcl
