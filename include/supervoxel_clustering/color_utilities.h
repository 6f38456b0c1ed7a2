/* color_utilities.h -- ColorUtilities with the reference's public surface, signature for signature
 * (reference: include/supervoxel_clustering/color_utilities.h:53-91, src/color_utilities.cpp).
 * mean_color / rgb2lab / lab_ciede00 / rgb_eucl run the device kernels of the f3ps path (colour.cuh) through the C ABI, so a
 * caller sees the numbers the merge loop uses.  As in the reference, the functions returning pointers hand back a
 * `new[]`-allocated array the caller owns (the reference never frees them, src/color_utilities.cpp:63,103,136). */
#ifndef F3PS_COLORUTILITIES_H_
#define F3PS_COLORUTILITIES_H_

#include "pcl_shim.h"

struct Color {
    uint8_t data[3];
};

typedef pcl::PointXYZRGBA PointT;
typedef pcl::Supervoxel<PointT> SupervoxelT;

const float RGB_RANGE = 441.672943;
const float LAB_RANGE = 137.3607;

class ColorUtilities {
    static float* color_conversion(float in[3], int code);          /* code: 0 = RGB -> L*a*b*, 1 = L*a*b* -> RGB */
    static float ciede00_test(float L1, float a1, float b1, float L2, float a2, float b2, float result);

    ColorUtilities() {}

public:
    /* colour of pcl::GlasbeyLUT at label % size.  PCL's table is not in /root/reference; this is a generated 256-entry
     * palette of maximally distinct colours, so label2color / color2label round-trip but the colours differ from PCL's */
    static uint8_t* get_glasbey(uint32_t label);
    /* running mean of the voxels' uint8 colours in voxels_ order (src/color_utilities.cpp:117-142) */
    static float* mean_color(SupervoxelT::Ptr s);
    static float* rgb2lab(float rgb[3]);
    static float* lab2rgb(float lab[3]);
    static float lab_ciede00(float lab1[3], float lab2[3], double kL = 1.0, double kC = 1.0, double kH = 1.0);
    static float rgb_eucl(float rgb1[3], float rgb2[3]);
    /* the reference's print-only self checks (src/color_utilities.cpp:324-499) */
    static void rgb_test();
    static void lab_test();
    static void convert_test();
};

#endif
