// oracle/ref_shim_png/opencv2/opencv.hpp -- stand-in for the OpenCV 2.4 header of the reference's
// depth-fusion/utils.hpp (utils.hpp:11) with a WORKING cv::imread for 16-bit grayscale PNGs, so that the reference's
// own `demo` (its main, its cudaMemcpys, its kernel, its PLY writer) can be built unchanged into oracle/_ref/demo_ref
// and timed end to end on the reference's sample sequence.  Only the PNG decode is substituted (OpenCV C++ is absent
// from this image): it goes through the engine library's zlib-based reader, which the tests pin against cv2.  The
// reference hard-codes 640 x 480 depth frames (demo.cu:87-88), and so does this shim.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#define CV_LOAD_IMAGE_UNCHANGED -1
extern "C" int tsdf_read_depth_png_u16(const char* path, int im_height, int im_width, unsigned short* depth_mm);
namespace cv {
struct Mat {
  int rows = 0, cols = 0;
  std::shared_ptr<std::vector<unsigned short>> data;
  bool empty() const { return !data || data->empty(); }
  template <class T> T at(int r, int c) const { return (T)(*data)[(size_t)r * cols + c]; }
};
inline Mat imread(const std::string& path, int) {
  Mat m;
  m.rows = 480;
  m.cols = 640;
  m.data = std::make_shared<std::vector<unsigned short>>((size_t)m.rows * m.cols);
  if (tsdf_read_depth_png_u16(path.c_str(), m.rows, m.cols, m.data->data()) != 0) m.data.reset();
  return m;
}
inline int waitKey(int) { return 0; }
}  // namespace cv
