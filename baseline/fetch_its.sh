#!/usr/bin/env bash
# TEST / BENCH INFRASTRUCTURE ONLY — stages the UNMODIFIED reference model files for the model-level harness
# (baseline/its_harness.py): ITS/models/*.py (the g2 model the training configs use) and the g4 variant's MIMOUNet.py
# (ITS/results_1mlp_g4/code, patch_size_global=4; its layers.py / vmamba_layers.py are byte-identical to ITS/models).
# Output goes to baseline/_ref/ only, which is git-ignored (reference sources never enter the history) but NOT
# gpurun-ignored, so the files travel to the GPU box where /root/reference does not exist.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
DST="$HERE/_ref/its_ref/models"
if [ ! -d "$REF/ITS/models" ]; then
    if [ -f "$DST/MIMOUNet.py" ]; then echo "[fetch_its] reference tree absent, using the staged copy in $DST"; exit 0; fi
    echo "[fetch_its] reference tree absent and nothing staged" >&2; exit 1
fi
mkdir -p "$DST"
for f in MIMOUNet.py layers.py vmamba_layers.py csm_triton.py; do cp -f "$REF/ITS/models/$f" "$DST/$f"; done
cmp -s "$REF/ITS/models/layers.py" "$REF/ITS/results_1mlp_g4/code/layers.py"
cmp -s "$REF/ITS/models/vmamba_layers.py" "$REF/ITS/results_1mlp_g4/code/vmamba_layers.py"
cp -f "$REF/ITS/results_1mlp_g4/code/MIMOUNet.py" "$DST/MIMOUNet_g4.py"
echo "[fetch_its] staged $(ls "$DST" | tr '\n' ' ')-> $DST"
