#!/bin/bash
# here: the product library and the -DWG_TIMING developer build (lib_timing/) side by side
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
ADMMB_BUILD_DIR=lib_timing ADMMB_EXTRA_NVCC_FLAGS=-DWG_TIMING python -c "
import importlib.util
spec = importlib.util.spec_from_file_location('b', 'admm-library_b200/build.py'); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True)"
