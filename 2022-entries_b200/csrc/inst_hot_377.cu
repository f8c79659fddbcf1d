#ifndef NO_MUL_BY_CALL
#define EC_MUL_BY_CALL 1
#endif
#include "launch_hot.inl"
namespace b200msm { template struct LaunchHot<Bls12_377>; }
