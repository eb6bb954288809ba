#!/usr/bin/env bash
# TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.
#
# Stages the UNMODIFIED reference (GANtastic3/MaskCycleGAN-VC, MIT) under oracle/_ref/ so that it
# travels to the GPU box with the gpurun snapshot (oracle/_ref/ is git-ignored: reference sources
# never enter this repository's history).  The reference is pure Python -- there is nothing to
# compile; "building" it is copying the five packages its train.py / test.py import:
#   mask_cyclegan_vc/{model,train,test,utils}.py  args/  dataset/  logger/  saver/
# Used by:  bench.py --impl reference and bench.py's cpu_baseline (kind "reference": the reference's
# own Generator / Discriminator modules and Adam on the host cores), tests/test_system_dropin.py
# (the unmodified train.py / test.py driven through the shim), tests/test_oracle.py (live pin).
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/mask_cyclegan_vc/model.py" ]; then
  echo "build_ref: $SRC does not hold the reference (no mask_cyclegan_vc/model.py)" >&2
  exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
for d in mask_cyclegan_vc args dataset logger saver; do
  cp -r "$SRC/$d" "$DST/$d"
done
cp "$SRC/LICENSE" "$DST/LICENSE"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find mask_cyclegan_vc args dataset logger saver -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "build_ref: staged $(find "$DST" -name '*.py' | wc -l) reference files under $DST"
