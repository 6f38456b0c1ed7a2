// pcd_io.h -- minimal PCD v0.7 reader/writer for the CLI (stands in for pcl::io::loadPCDFile,
// src/supervoxel_clustering.cpp:313): ascii / binary / binary_compressed, fields x y z rgb|rgba [label].
#pragma once
#include <string>
#include "supervoxel_clustering/pcl_shim.h"

namespace f3ps {
// returns 0 on success, -1 on failure (the reference ignores the return value and carries on with an empty cloud)
int loadPCDFile(const std::string& path, pcl::PointCloud<pcl::PointXYZRGBL>& cloud);
int savePCDFileASCII(const std::string& path, const pcl::PointCloud<pcl::PointXYZL>& cloud);
}
