#define EC_MUL_BY_CALL 1
#include "launch_cold.inl"
namespace b200msm { template struct LaunchCold<Bls12_377_sw>; }
