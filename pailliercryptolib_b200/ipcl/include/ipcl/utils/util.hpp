// util.hpp -- error convention of the ipcl:: layer.
// Same contract as the reference (ipcl/include/ipcl/utils/util.hpp:23-34):
// ERROR_CHECK(cond, msg) throws std::runtime_error whose text is
// "\nFile: <file>\nLine: <line>\nError: <msg>".  The OpenMP thread budgeting
// and CPU-feature probing of the reference are x86 host scheduling and have no
// counterpart here (the batch goes to the GPU in one submission).
#ifndef IPCL_B200_UTILS_UTIL_HPP_
#define IPCL_B200_UTILS_UTIL_HPP_

#include <sstream>
#include <stdexcept>
#include <string>

#include "ipcl/utils/common.hpp"

namespace ipcl {

inline std::string build_log(const char* file, int line, const std::string& msg) {
  std::ostringstream log;
  log << "\nFile: " << file << "\nLine: " << line << "\nError: " << msg;
  return log.str();
}

#define ERROR_CHECK(e, ...)                                                 \
  do {                                                                      \
    if (!(e))                                                               \
      throw std::runtime_error(                                             \
          ::ipcl::build_log(__FILE__, __LINE__, __VA_ARGS__));              \
  } while (0)

// turns a non-zero status of the C ABI (include/ipcl_b200.h) into the same
// exception shape, the way mod_exp.cpp:518-523 turns mbx lane status into one
void check_device_status(int rc, const char* what, const char* file, int line);
#define DEVICE_CHECK(call) \
  ::ipcl::check_device_status((call), #call, __FILE__, __LINE__)

}  // namespace ipcl
#endif  // IPCL_B200_UTILS_UTIL_HPP_
