// STAND-IN (test infrastructure): see gtsam/geometry/Pose3.h
#include "../geometry/Pose3.h"
