#include "launch_cold.inl"
namespace b200msm { template struct LaunchCold<Bls12_377>; }
