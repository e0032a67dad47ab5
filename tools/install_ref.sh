#!/bin/bash
# Installs the UNMODIFIED reference model classes next to the repo for the CPU arm of bench.py
# (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`).
#
# The reference (wadaniel/marlpde) is a set of Python scripts without packaging metadata, so there is nothing for
# `pip install --target baseline/_ref /root/reference` to install; the equivalent is a verbatim copy of its
# python/_model directory into the git-ignored (NOT gpurun-ignored) baseline/_ref/, which travels to the GPU box.
# Nothing under baseline/_ref/ is committed, imported by the product (marlpde_b200/) or read by the tests.
#
# usage: bash tools/install_ref.sh            (build container only: needs /root/reference)
set -eu
here=$(cd "$(dirname "$0")/.." && pwd)
src=/root/reference/python/_model
dst=$here/baseline/_ref/_model
if [ ! -d "$src" ]; then
    echo "install_ref: $src not found (not the build container) -- keeping whatever $dst holds" >&2
    exit 0
fi
mkdir -p "$dst"
cp -f "$src"/Burger.py "$src"/KS.py "$src"/Diffusion.py "$src"/Advection.py "$dst"/
( cd "$src" && sha256sum Burger.py KS.py Diffusion.py Advection.py ) > "$dst"/SHA256SUMS
echo "install_ref: copied $(ls "$dst" | wc -l) files to $dst"
