#!/bin/bash
# Install the unmodified reference into baseline/_ref (git-ignored; travels to the GPU box).
# `pip install /root/reference` is impossible: the reference has no setup.py/pyproject.toml
# ("Directory '/root/reference' is not installable"), it is five top-level scripts — so the
# scripts are copied verbatim (no edits) and imported from there by `bench.py --impl reference`.
set -e
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference \
  || cp /root/reference/*.py baseline/_ref/
ls baseline/_ref
