#!/usr/bin/env bash
# Compiles the shim test against the stub Eigen/PCL headers (host-only; running it needs a B200).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="${HERE}/../.."
/usr/bin/g++ -std=c++14 -O2 -Wall -I"${HERE}/stubs" -I"${REPO}/include" -I"${REPO}/go-rio_b200/include" \
  "${HERE}/test_shim.cpp" -o "${HERE}/test_shim" -L"${REPO}/go-rio_b200" -lapdgicp -Wl,-rpath,'$ORIGIN/..'
echo "built ${HERE}/test_shim"
