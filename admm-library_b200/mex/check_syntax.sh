#!/usr/bin/env bash
# Syntax check of the MEX gateway against the stub mex.h (MATLAB is absent from this image).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
cxx=/usr/bin/g++; [ -x "$cxx" ] || cxx=g++
"$cxx" -std=c++17 -fsyntax-only -Wall -Wextra -I"$here/stub" -I"$here/../../include" "$here/admm_mex.cpp"
