// Stand-in for pcl/console/print.h: silent (oracle/ref_shim/README.md).  Test infrastructure only.
#pragma once
namespace pcl { namespace console {
inline void print_debug(const char*, ...) {} inline void print_info(const char*, ...) {} inline void print_warn(const char*, ...) {}
inline void print_error(const char*, ...) {} inline void print_highlight(const char*, ...) {}
} }
