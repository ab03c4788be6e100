// STAND-IN (test infrastructure only), see pcl/point_types.h
#pragma once
#include <pcl/point_types.h>
