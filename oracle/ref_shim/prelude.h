// prelude.h — force-included (-include) in front of every unit of the reference build: all standard headers first, THEN
// the access specifiers are opened so that the test harness can seed the engine (the reference seeds from
// std::random_device), install map layers without OpenCV, and read the particles back.  The reference's source files
// themselves are compiled unmodified from /root/reference.
#pragma once
#include <math.h>
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <execution>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>
#include <opencv2/ml/ml.hpp>
#include <pcl/point_types.h>
#include <ros/ros.h>
#define private public
#define protected public
