// oracle/ref_shim/opencv2/opencv.hpp -- stand-in for the OpenCV 2.4 header the reference's
// depth-fusion/utils.hpp includes (utils.hpp:11).  OpenCV C++ is absent from this image and
// only ReadDepth (utils.hpp:74-86) touches it; the pinned builds feed depth from memory, so
// imread returns an empty Mat.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#define CV_LOAD_IMAGE_UNCHANGED -1
namespace cv {
struct Mat {
  bool empty() const { return true; }
  template <class T> T at(int, int) const { return T(); }
};
inline Mat imread(const std::string&, int) { return Mat(); }
inline int waitKey(int) { return 0; }
}  // namespace cv
