// pcd_dump -- reads a PCD with the CLI's reader (pcd_io.cpp) and prints what it decoded:
//   "<points> <width> <height> <finite xyz> <fnv1a64 of x,y,z,rgba,label bits>"
// tests/test_pcd_io.py compares this with the Python reader on ascii / binary / binary_compressed files.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "pcd_io.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: pcd_dump file.pcd\n"); return 2; }
    pcl::PointCloud<pcl::PointXYZRGBL> cloud;
    if (f3ps::loadPCDFile(argv[1], cloud) != 0) { printf("error\n"); return 1; }
    uint64_t h = 1469598103934665603ull; size_t finite = 0;
    auto mix = [&](uint32_t v) { for (int i = 0; i < 4; ++i) { h ^= (v >> (8 * i)) & 255u; h *= 1099511628211ull; } };
    for (const auto& p : cloud.points) {
        uint32_t b[3]; memcpy(&b[0], &p.x, 4); memcpy(&b[1], &p.y, 4); memcpy(&b[2], &p.z, 4);
        if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) ++finite;
        for (int k = 0; k < 3; ++k) mix(std::isnan(k == 0 ? p.x : k == 1 ? p.y : p.z) ? 0x7fc00000u : b[k]);
        mix(p.rgba); mix(p.label);
    }
    printf("%zu %u %u %zu %016llx\n", cloud.size(), cloud.width, cloud.height, finite, (unsigned long long)h);
    return 0;
}
