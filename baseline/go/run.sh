#!/bin/bash
# Time the unmodified Go reference on a bench configuration (needs a Go toolchain >= 1.24 and a checkout of the reference).
# usage: baseline/go/run.sh <path to a checkout of JoshElkind/concurrent-raytracer-go> [gort_baseline flags]
set -e
REF=${1:?path to the reference checkout}; shift
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$REF/cmd/gort_baseline"
cp "$HERE/main.go" "$REF/cmd/gort_baseline/main.go"
cd "$REF" && go run ./cmd/gort_baseline "$@"
