// oracle/shim_handeye: TEST INFRASTRUCTURE (see se3.hpp): nothing of Sophus::SO3 is used on this path.
#pragma once
