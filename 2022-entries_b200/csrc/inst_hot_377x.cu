#define EC_MUL_BY_CALL 1
#include "launch_hot.inl"
namespace b200msm { template struct LaunchHot<Bls12_377_sw>; }
