// Driver of tests/test_resize_tables.py: prints the host-side cv::resize tables (fealess_b200/csrc/resize_tables.h), no GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../fealess_b200/csrc/resize_tables.h"
int main(int argc, char** argv) {
  if (argc != 4) { std::printf("usage: resize_tables_test <src> <dst> <clamp 0|1>\n"); return 2; }
  std::vector<int> ofs; std::vector<float> w; std::vector<short> iw;
  fl_resize_axis(std::atoi(argv[1]), std::atoi(argv[2]), std::atoi(argv[3]) != 0, ofs, w, iw);
  for (size_t d = 0; d < ofs.size(); ++d) {
    unsigned a, b; std::memcpy(&a, &w[2 * d], 4); std::memcpy(&b, &w[2 * d + 1], 4);
    std::printf("%d %08x %08x %d %d\n", ofs[d], a, b, (int)iw[2 * d], (int)iw[2 * d + 1]);
  }
  return 0;
}
