// oracle/ref_shim/pcl/io/pcd_io.h -- TEST INFRASTRUCTURE: stand-in so that the reference's Algorithms.hpp and FileRoutines.hpp
// compile in this PCL-less container.  FileRoutines.hpp::readPointCloud names loadPCDFile / PCL_ERROR and, through PCL's
// own includes, boost::split / boost::is_any_of (readCameraLocations, FileRoutines.hpp:84): loadPCDFile always fails here
// (nothing calls readPointCloud); boost::split is the plain "split on any of these characters, keep empty fields" it is
// in Boost (token_compress_off), which is all the pose-file reader needs.
#pragma once
#include <string>
#include <vector>
#define PCL_ERROR(...) ((void)0)
namespace pcl { namespace io {
template <typename PointT, typename CloudT> int loadPCDFile(const std::string&, CloudT&) { return -1; }
} }
namespace boost {
struct is_any_of_t { std::string chars; };
inline is_any_of_t is_any_of(const char* s) { return is_any_of_t{s}; }
inline void split(std::vector<std::string>& out, const std::string& in, const is_any_of_t& pred) {
    out.clear();
    std::string cur;
    for (char ch : in) {
        if (pred.chars.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur.push_back(ch);
    }
    out.push_back(cur);
}
}
