// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in; kd_tree_two.h includes
// this ROS header but uses nothing from it.
#pragma once
