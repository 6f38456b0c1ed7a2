/* color_utilities.h -- ColorUtilities with the reference's public surface
 * (reference: include/supervoxel_clustering/color_utilities.h:53-91, src/color_utilities.cpp).
 * Every function runs the device kernels of the f3ps path (colour.cuh) through the C ABI; results
 * are returned by value-owning arrays instead of leaked new[] buffers. */
#ifndef F3PS_COLORUTILITIES_H_
#define F3PS_COLORUTILITIES_H_

#include <array>
#include "pcl_shim.h"

typedef pcl::PointXYZRGBA PointT;
typedef pcl::Supervoxel<PointT> SupervoxelT;

const float RGB_RANGE = 441.672943f;
const float LAB_RANGE = 137.3607f;

class ColorUtilities {
    ColorUtilities() {}
public:
    /* 256-entry distinct-colour palette indexed by label % 256 (stands in for pcl::GlasbeyLUT, viewer only) */
    static std::array<uint8_t, 3> get_glasbey(uint32_t label);
    /* running mean of the voxels' uint8 colours in voxels_ order (src/color_utilities.cpp:117-142) */
    static std::array<float, 3> mean_color(SupervoxelT::Ptr s);
    static std::array<float, 3> rgb2lab(const float rgb[3]);
    static float lab_ciede00(const float lab1[3], const float lab2[3]);
    static float rgb_eucl(const float rgb1[3], const float rgb2[3]);
    /* the reference's print-only self checks, here returning the maximum error */
    static float rgb_test();
    static float lab_test();
};

#endif
