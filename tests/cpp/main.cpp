// test driver; mirrors /root/reference/test/main.cpp:9-45
#include <string>

#include "check.hpp"
#include "ipcl/ipcl.hpp"

int main(int argc, char** argv) {
  std::string filter = argc > 1 ? argv[1] : "";
  ipcl::initializeContext("GPU");
  int rc = check::run_all(filter);
  ipcl::terminateContext();
  return rc ? 1 : 0;
}
