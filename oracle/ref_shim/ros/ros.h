// ros/ros.h — STAND-IN (test infrastructure only): the reference uses ROS in these units for log lines alone.
#pragma once
#include <iostream>
#include <sstream>
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_DEBUG_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_DEBUG(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
