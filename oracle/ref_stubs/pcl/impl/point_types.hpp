// stand-in for pcl::PointXYZ (16-byte {x, y, z, pad}); see ../../Eigen/Dense for why this exists
#pragma once
#include <cmath>
#include <math.h>
#include "../../Eigen/Dense"
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad; };
}
