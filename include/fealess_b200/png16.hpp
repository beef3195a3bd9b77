// fealess_b200/png16.hpp - decoder for the rendered template depth images of a FEALESS feature directory
// (<path>/depth/<template_id>.png, 16-bit grey, read by the reference with cv::imread(path, -1) at
// CadReco/obj_reco_lmicp.cpp:156-157).  Hosts that build with the OpenCV SDK use cv::imread; this image has no OpenCV C++ SDK,
// so the C++ CObjRecoLmICP mirror reads the files itself: non-interlaced greyscale PNG, 8 or 16 bits per sample, zlib for the
// inflate step (-lz).  tests/test_cpp_reco.py compares it with cv2.imread on files written by cv2.imwrite.
#ifndef FEALESS_B200_PNG16_HPP
#define FEALESS_B200_PNG16_HPP

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace fealess_b200 {

// Reads `path` into `pix` (row-major, width*height samples, 8-bit samples widened).  false: unreadable / unsupported file.
inline bool read_png_gray16(const std::string& path, std::vector<uint16_t>& pix, int& width, int& height) {
  pix.clear(); width = height = 0;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> file;
  unsigned char buf[65536];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + got);
  std::fclose(f);
  static const unsigned char SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 + 25 || std::memcmp(file.data(), SIG, 8) != 0) return false;
  auto be32 = [&](size_t p) { return ((uint32_t)file[p] << 24) | ((uint32_t)file[p + 1] << 16) | ((uint32_t)file[p + 2] << 8) | file[p + 3]; };
  int depth = 0;
  std::vector<unsigned char> idat;
  for (size_t p = 8; p + 12 <= file.size();) {
    const uint32_t len = be32(p);
    if (p + 12 + (size_t)len > file.size()) return false;
    const char* type = (const char*)&file[p + 4];
    const size_t body = p + 8;
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len < 13) return false;
      width = (int)be32(body); height = (int)be32(body + 4); depth = file[body + 8];
      const int color = file[body + 9], interlace = file[body + 12];
      if (width <= 0 || height <= 0 || color != 0 || interlace != 0 || (depth != 8 && depth != 16)) return false;
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), file.begin() + body, file.begin() + body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    p += 12 + (size_t)len;
  }
  if (!depth || idat.empty()) return false;
  const size_t bpp = (size_t)depth / 8, stride = (size_t)width * bpp;
  std::vector<unsigned char> raw((stride + 1) * (size_t)height);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return false;
  // undo the per-row filters (PNG specification, section 9: None, Sub, Up, Average, Paeth) in place
  std::vector<unsigned char> prev(stride, 0);
  pix.resize((size_t)width * height);
  for (int y = 0; y < height; ++y) {
    unsigned char* row = &raw[(stride + 1) * (size_t)y];
    const int filter = row[0];
    unsigned char* cur = row + 1;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
      int add = 0;
      switch (filter) {
        case 0: break;
        case 1: add = a; break;
        case 2: add = b; break;
        case 3: add = (a + b) >> 1; break;
        case 4: { const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c); add = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: return false;
      }
      cur[i] = (unsigned char)(cur[i] + add);
    }
    std::memcpy(prev.data(), cur, stride);
    uint16_t* out = &pix[(size_t)y * width];
    if (depth == 16) for (int x = 0; x < width; ++x) out[x] = (uint16_t)((cur[2 * x] << 8) | cur[2 * x + 1]);   // samples are big-endian
    else for (int x = 0; x < width; ++x) out[x] = cur[x];
  }
  return true;
}

}  // namespace fealess_b200

#endif  // FEALESS_B200_PNG16_HPP
