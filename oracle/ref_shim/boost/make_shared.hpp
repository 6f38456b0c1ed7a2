// Stand-in: boost::shared_ptr / make_shared = the std ones (oracle/ref_shim/README.md).  Test infrastructure only.
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; using std::make_shared; }
