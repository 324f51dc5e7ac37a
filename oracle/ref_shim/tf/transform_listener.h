// ros_stub: see ros/ros.h
#include "../ros/ros.h"
