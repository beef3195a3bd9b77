// Driver of tests/test_cpp_reco.py: the C++ CObjRecoCAD / CObjRecoLmICP mirror (include/fealess_b200/obj_reco.hpp).
//   reco_test png <file>                 decode a grey PNG: "w h fnv1a64" of the samples            (no GPU)
//   reco_test addobj <dir>               AddObj status, classes, model depth images, hash of image 0 (no GPU)
//   reco_test badparams <dir>            Recognition argument checks                                 (no GPU)
//   reco_test run <dir> <frame.bin>...   Create -> AddObj -> Recognition per frame; frame.bin = int32 W, H, double fx, fy, cx, cy, BGR bytes, depth u16
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fealess_b200/obj_reco.hpp"

static unsigned long long fnv(const void* p, size_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= ((const unsigned char*)p)[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  try {
    if (argc == 3 && !std::strcmp(argv[1], "png")) {
      std::vector<uint16_t> pix; int w = 0, h = 0;
      if (!fealess_b200::read_png_gray16(argv[2], pix, w, h)) { std::printf("unreadable\n"); return 1; }
      std::printf("%d %d %016llx\n", w, h, fnv(pix.data(), pix.size() * 2));
      return 0;
    }
    if (argc == 3 && !std::strcmp(argv[1], "addobj")) {
      CObjRecoLmICP reco;
      const int rc = reco.AddObj(argv[2]);
      std::printf("status %08x\n", (unsigned)rc);
      if (rc == 0) {
        const CObjRecoLmICP::ModelDepth* md = reco.modelDepth(0);
        std::printf("classes %d depths %zu", reco.detector()->numClasses(), reco.numModelDepths());
        if (md) std::printf(" %d %d %016llx", md->width, md->height, fnv(md->mm.data(), md->mm.size() * 2));
        std::printf("\n");
      }
      return 0;
    }
    if (argc == 3 && !std::strcmp(argv[1], "badparams")) {
      CObjRecoCAD* reco = CObjRecoCAD::Create();
      std::vector<unsigned char> rgb(640 * 480 * 3); std::vector<unsigned short> dep(640 * 480);
      TImageU tRGB = {0.0, rgb.data(), 640, 480}; TImageU16 tDepth = {0.0, dep.data(), 640, 480};
      TCamIntrinsicParam K = {640, 480, 608., 608., 320., 240., std::vector<double>()};
      std::vector<TObjRecoResult> res;
      std::printf("no object %08x\n", (unsigned)reco->Recognition(tRGB, tDepth, K, res));
      std::printf("addobj %08x\n", (unsigned)reco->AddObj(argv[2]));
      TImageU16 small = tDepth; small.nHeight = 100;
      std::printf("depth size %08x\n", (unsigned)reco->Recognition(tRGB, small, K, res));
      TCamIntrinsicParam K2 = K; K2.nWidth = 320;
      std::printf("intrinsics size %08x\n", (unsigned)reco->Recognition(tRGB, tDepth, K2, res));
      TImageU null_img = tRGB; null_img.pData = nullptr;
      std::printf("null image %08x\n", (unsigned)reco->Recognition(null_img, tDepth, K, res));
      TImageU neg = tRGB; neg.dTimestamp = -1.0;
      std::printf("negative timestamp %08x\n", (unsigned)reco->Recognition(neg, tDepth, K, res));
      std::printf("unsupported type %s\n", CObjRecoCAD::Create(CObjRecoCAD::EObjReco_BB8) ? "object" : "null");
      std::printf("misc %d %d %d\n", reco->ClearObj(), reco->SetROI(tRGB), reco->Train("", TScanPackage(), TTrainParam()));
      CObjRecoCAD::Destroy(reco);
      return 0;
    }
    // run <dir> <frame.bin>...            the reference's behaviour: ICP of matches[0]
    // hyp <dir> <top_k> <per_class> <th_obj_dist> <frame.bin>...   SetHypotheses: top-K / per-class selection + nonMaximumSuppression
    const bool hyp_mode = argc >= 7 && !std::strcmp(argv[1], "hyp");
    if ((argc >= 4 && !std::strcmp(argv[1], "run")) || hyp_mode) {
      CObjRecoCAD* reco = CObjRecoCAD::Create(CObjRecoCAD::EObjReco_LmICP);
      std::printf("addobj %08x\n", (unsigned)reco->AddObj(argv[2]));
      if (hyp_mode) static_cast<CObjRecoLmICP*>(reco)->SetHypotheses(std::atoi(argv[3]), std::atoi(argv[4]) != 0, (float)std::atof(argv[5]));
      for (int a = hyp_mode ? 6 : 3; a < argc; ++a) {
        FILE* f = std::fopen(argv[a], "rb");
        if (!f) { std::printf("cannot open %s\n", argv[a]); return 2; }
        int wh[2]; double k[4];
        if (std::fread(wh, 4, 2, f) != 2 || std::fread(k, 8, 4, f) != 4) return 2;
        std::vector<unsigned char> rgb((size_t)wh[0] * wh[1] * 3); std::vector<unsigned short> dep((size_t)wh[0] * wh[1]);
        if (std::fread(rgb.data(), 1, rgb.size(), f) != rgb.size() || std::fread(dep.data(), 2, dep.size(), f) != dep.size()) return 2;
        std::fclose(f);
        TImageU tRGB = {1.0, rgb.data(), wh[0], wh[1]}; TImageU16 tDepth = {1.0, dep.data(), wh[0], wh[1]};
        TCamIntrinsicParam K = {wh[0], wh[1], k[0], k[1], k[2], k[3], std::vector<double>()};
        std::vector<TObjRecoResult> res;
        for (int rep = 0; rep < 2; ++rep) {                        // twice: the second call finds the crops already on the device
          const int rc = reco->Recognition(tRGB, tDepth, K, res);
          std::printf("frame %d status %08x results %zu", a - (hyp_mode ? 6 : 3), (unsigned)rc, res.size());
          for (size_t i = 0; i < res.size(); ++i) {
            std::printf(" %s", res[i].strObjTag.c_str());
            for (int j = 0; j < 16; ++j) { unsigned u; std::memcpy(&u, &res[i].tWorld2Cam[j], 4); std::printf(" %08x", u); }
          }
          std::printf("\n");
        }
      }
      CObjRecoCAD::Destroy(reco);
      return 0;
    }
  } catch (const cv::Exception& e) {
    std::printf("cv::Exception: %s\n", e.what());
    return 3;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 4;
  }
  std::printf("usage: reco_test png <file> | addobj <dir> | badparams <dir> | run <dir> <frame.bin>... | hyp <dir> <top_k> <per_class> <th_obj_dist> <frame.bin>...\n");
  return 2;
}
