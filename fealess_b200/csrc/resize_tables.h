// resize_tables.h - host side of the device rescale (resize.cu): the per-axis index / weight tables of cv::resize INTER_LINEAR,
// computed like OpenCV's own table loops (imgproc/src/resize.cpp: double scale, float fractional part; columns clamp index and
// weight at the borders, rows only have their index clipped when fetched; 8U weights = saturate_cast<short>(w * 2048)).
// Plain C++ so that tests/test_resize_tables.py can compile it without CUDA and compare with oracle/resize_oracle.py.
#ifndef FL_RESIZE_TABLES_H
#define FL_RESIZE_TABLES_H
#include <cmath>
#include <vector>

inline void fl_resize_axis(int ssize, int dsize, bool clamp_weights, std::vector<int>& ofs, std::vector<float>& w, std::vector<short>& iw) {
  const double inv_scale = (double)dsize / ssize, scale = 1.0 / inv_scale;
  ofs.resize((size_t)dsize); w.resize((size_t)dsize * 2); iw.resize((size_t)dsize * 2);
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= (float)s;
    if (clamp_weights) {
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    ofs[d] = s; w[2 * d] = 1.f - f; w[2 * d + 1] = f;
    iw[2 * d] = (short)std::lrintf(w[2 * d] * 2048.f); iw[2 * d + 1] = (short)std::lrintf(w[2 * d + 1] * 2048.f);
  }
}
#endif
