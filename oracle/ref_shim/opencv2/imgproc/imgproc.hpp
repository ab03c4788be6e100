// STAND-IN (test infrastructure only), see opencv2/core/core.hpp
#pragma once
#include <opencv2/core/core.hpp>
