#!/usr/bin/env python3
"""Fixture import script (run in the build container only; /root/reference is absent on the GPU box).

Copies the arkworks-generated BLS12-377 G1 MSM known-answer vectors that the reference ships in its
prize-7 mobile harness into tests/golden/ and records their sha256:

  /root/reference/open-division/prize7-msm-mobile/gnark/msm/bls12-377/{points,scalars,result}
      written by gnark/src/lib.rs:245-283 (result = ark_ec::msm::VariableBaseMSM::multi_scalar_mul,
      ark 0.3, CanonicalSerialize) and read back by gnark/msm/bls12-377/marshal_zprize.go:24-150.
      5 instances, N = 16384, 16384, 256, 16, 65536.
  /root/reference/open-division/prize7-msm-mobile/yyteam/mobileMSM/app/src/main/assets/zero-points-test/
      {points,scalars}: 256 points all flagged infinity (edge case; expected result = infinity).

These are DATA fixtures produced by the reference's own oracle (arkworks), not reference sources.
Format: repeat { u64 LE n ; n x 48-byte compressed G1Affine } for points,
repeat { u64 LE n ; n x 32-byte LE canonical scalar } for scalars, one 48-byte compressed point per
instance for result.
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference/open-division/prize7-msm-mobile"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "ark377_points.bin": f"{REF}/gnark/msm/bls12-377/points",
    "ark377_scalars.bin": f"{REF}/gnark/msm/bls12-377/scalars",
    "ark377_result.bin": f"{REF}/gnark/msm/bls12-377/result",
    "ark377_allinf_points.bin": f"{REF}/yyteam/mobileMSM/app/src/main/assets/zero-points-test/points",
    "ark377_allinf_scalars.bin": f"{REF}/yyteam/mobileMSM/app/src/main/assets/zero-points-test/scalars",
}


def main():
    manifest = {}
    for dst, src in FILES.items():
        out = os.path.join(HERE, dst)
        shutil.copyfile(src, out)
        os.chmod(out, 0o644)
        with open(out, "rb") as f:
            data = f.read()
        manifest[dst] = {"source": src.replace("/root/reference/", ""), "bytes": len(data),
                         "sha256": hashlib.sha256(data).hexdigest()}
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
