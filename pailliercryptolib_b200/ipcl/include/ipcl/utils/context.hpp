// context.hpp -- runtime context (ipcl/include/ipcl/utils/context.hpp:25-44).
// initializeContext accepts the reference's choices ("DEFAULT", "CPU", "QAT",
// "HYBRID", any case) plus "GPU"/"B200"; every one of them brings up the CUDA
// back-end, because it is the only back-end.  isQATRunning/isQATActive are
// kept and return false.
#ifndef IPCL_B200_UTILS_CONTEXT_HPP_
#define IPCL_B200_UTILS_CONTEXT_HPP_

#include <string>

namespace ipcl {

bool initializeContext(const std::string runtime_choice);
bool terminateContext(void);
bool isQATRunning(void);
bool isQATActive(void);
// true once the sm_100a device context is up
bool isGPURunning(void);

}  // namespace ipcl
#endif  // IPCL_B200_UTILS_CONTEXT_HPP_
