// Stand-in for pcl/point_types.h: the one point type src/testing.cpp uses (oracle/ref_shim/README.md).  Test infrastructure only.
#pragma once
#include <cstdint>
namespace pcl { struct PointXYZL { float x, y, z; uint32_t label; }; }
