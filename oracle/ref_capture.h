// ref_capture.h -- globals shared by ref_standins.cpp and ref_clustering_wrap.cpp (test infrastructure)
#pragma once
#include <cstdint>
#include <vector>
namespace f3ps_ref {
struct MergeLine { uint32_t edges_left, regions_left, a, b; float w; };
extern const int16_t* lab_lut;                 // OpenCV's 33^3 Lab table (the oracle's, pinned against cv2 4.13.0)
extern std::vector<MergeLine>* sink;           // where print_debug("left: ...") lines go
}
