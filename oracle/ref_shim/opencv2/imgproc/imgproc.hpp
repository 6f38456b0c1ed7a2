// Stand-in: see f3ps_ref_standins.h (oracle/ref_shim/README.md).  Test infrastructure only.
#pragma once
#include "../../f3ps_ref_standins.h"
