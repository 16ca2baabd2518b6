// TEST INFRASTRUCTURE — empty stand-in for <ros/ros.h>.  The reference's visualization.cpp includes it (visualization.cpp:30)
// without using anything from it; this lets the unmodified file compile into oracle/_ref/libref_pose.so.  Not ROS code.
#pragma once
