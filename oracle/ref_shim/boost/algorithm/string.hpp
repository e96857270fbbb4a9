// oracle/ref_shim/boost/algorithm/string.hpp -- TEST INFRASTRUCTURE: stand-in (the reference's drivers say
// `using namespace boost::algorithm;`); boost::split / is_any_of live in pcl/io/pcd_io.h of this shim.
#pragma once
#include <pcl/io/pcd_io.h>
namespace boost { namespace algorithm {} }
