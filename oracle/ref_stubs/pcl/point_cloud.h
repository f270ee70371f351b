// stand-in for pcl::PointCloud<T> (a `points` vector behind a shared pointer)
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <class P> struct PointCloud {
  std::vector<P> points;
  typedef std::shared_ptr<PointCloud<P>> Ptr;
};
}
