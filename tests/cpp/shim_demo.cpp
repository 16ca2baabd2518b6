// Exercises the C++ class shim (include/monocular_pose_estimator_b200/shim.h, stand-in types) the way MPENode uses the
// reference library: configure the public fields, setMarkerPositions, then estimateBodyPose per frame.
// Input: a little binary scene file written by tests/test_gpu_cpp_shim.py.  Output: one text line per frame.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#define MPE_SHIM_FORCE_STANDIN 1
#include "monocular_pose_estimator_b200/shim.h"

using namespace monocular_pose_estimator;

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: shim_demo scene.bin [stage]\n"); return 2; }
  const bool stage_mode = argc > 2 && std::string(argv[2]) == "stage";   // stage by stage on the host instead of the device loop
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int hdr[4];
  if (fread(hdr, sizeof(int), 4, f) != 4) return 2;
  const int n_frames = hdr[0], w = hdr[1], h = hdr[2], n_leds = hdr[3];
  double K[9], D[5], params[11];
  std::vector<double> markers(3 * n_leds), times(n_frames);
  if (fread(K, 8, 9, f) != 9 || fread(D, 8, 5, f) != 5 || fread(markers.data(), 8, markers.size(), f) != markers.size() ||
      fread(params, 8, 11, f) != 11 || fread(times.data(), 8, times.size(), f) != times.size()) return 2;
  std::vector<uint8_t> frames((size_t)n_frames * w * h);
  if (fread(frames.data(), 1, frames.size(), f) != frames.size()) return 2;
  fclose(f);

  try {
    PoseEstimator pe;
    pe.setDeviceLoop(!stage_mode);
    for (int i = 0; i < 9; ++i) pe.camera_matrix_K_.k[i] = K[i];                 // MPENode::cameraInfoCallback
    pe.camera_distortion_coeffs_.assign(D, D + 5);
    pe.detection_threshold_value_ = (int)params[0];                              // MPENode::dynamicParametersCallback
    pe.gaussian_sigma_ = params[1]; pe.min_blob_area_ = params[2]; pe.max_blob_area_ = params[3];
    pe.max_width_height_distortion_ = params[4]; pe.max_circular_distortion_ = params[5];
    pe.setBackProjectionPixelTolerance(params[6]); pe.setNearestNeighbourPixelTolerance(params[7]);
    pe.setCertaintyThreshold(params[8]); pe.setValidCorrespondenceThreshold(params[9]);
    pe.roi_border_thickness_ = (unsigned)params[10];
    List4DPoints pts; pts.resize(n_leds);
    for (int i = 0; i < n_leds; ++i) { pts(i)(0) = markers[3 * i]; pts(i)(1) = markers[3 * i + 1]; pts(i)(2) = markers[3 * i + 2]; pts(i)(3) = 1; }
    pe.setMarkerPositions(pts);
    for (int fi = 0; fi < n_frames; ++fi) {
      ImageT img; img.data = frames.data() + (size_t)fi * w * h; img.rows = h; img.cols = w; img.step = w;
      bool ok = pe.estimateBodyPose(img, times[fi]);
      const RectT& r = pe.regionOfInterest();
      printf("%d %d %d %d %d %d %u", fi, ok ? 1 : 0, r.x, r.y, r.width, r.height, pe.lastGaussNewtonIterations());
      Matrix4dT T = pe.getPredictedPose();
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) printf(" %.17g", T(a, b));
      printf("\n");
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
