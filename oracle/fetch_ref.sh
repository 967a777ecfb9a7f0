#!/bin/sh
# TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.
#
# Copies the UNMODIFIED reference scripts (skoohy/GPT-PINN: Klein-Gordon/, Burgers/, Allen-Cahn/,
# top-level *.py + AC.mat; not original/, not fig/) from the read-only checkout into
# oracle/_ref/, byte for byte.  oracle/_ref/ is git-ignored (no reference source enters the
# history) but NOT gpurun-ignored, so the copy travels to the GPU box, where
#   * bench.py --impl reference times the reference's own offline_generation / gpt_test on the
#     host cores (cpu_baseline.kind = "reference") and, with a GPU, the same modules on cuda;
#   * tools/run_reference_main.py runs oracle/_ref/Klein-Gordon/KG_main.py and Burgers/B_main.py
#     UNCHANGED with gpt-pinn_b200/dropin first on sys.path.
# A sha256 manifest is written next to the copy; tests compare it with /root/reference when both exist.
set -eu
SRC="${GPTPINN_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/Klein-Gordon" ]; then
    if [ -d "$DST/Klein-Gordon" ]; then
        echo "fetch_ref: $SRC absent, keeping the existing copy in $DST"
        exit 0
    fi
    echo "fetch_ref: $SRC absent and no copy in $DST" >&2
    exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
for d in Klein-Gordon Burgers Allen-Cahn; do
    mkdir -p "$DST/$d"
    for f in "$SRC/$d"/*.py "$SRC/$d"/*.mat; do
        [ -f "$f" ] && cp "$f" "$DST/$d/"
    done
done
(cd "$DST" && find . -type f ! -name MANIFEST.sha256 | sort | xargs sha256sum > MANIFEST.sha256)
echo "fetch_ref: copied $(find "$DST" -type f -name '*.py' | wc -l) reference scripts to $DST"
