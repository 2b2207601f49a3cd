#!/bin/bash
# Installs the unmodified reference (clipa_torch: open_clip + training) into baseline/_ref (git-ignored,
# shipped to the GPU box by gpurun).  /root/reference is read-only, so the wheel is built from a copy.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/clipa_refsrc baseline/_ref
cp -r /root/reference/clipa_torch /tmp/clipa_refsrc
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref /tmp/clipa_refsrc
ls baseline/_ref
# setup.py declares no package_data, so the wheel drops the data files the package loads at import time
# (BPE vocabulary, JSON model configs): add them next to the installed modules.
cp /root/reference/clipa_torch/open_clip/bpe_simple_vocab_16e6.txt.gz baseline/_ref/open_clip/
mkdir -p baseline/_ref/open_clip/model_configs
cp /root/reference/clipa_torch/open_clip/model_configs/*.json baseline/_ref/open_clip/model_configs/
du -sh baseline/_ref
