#!/bin/bash
# A/B helper: run a command with an older build of the library swapped in (tools/ab/libpsoap_old.so), then restore.
set -e
cd "$(dirname "$0")/.."
cp psoap_b200/csrc/libpsoap_b200.so /tmp/libpsoap_new.so
cp tools/ab/libpsoap_old.so psoap_b200/csrc/libpsoap_b200.so
"$@" || true
cp /tmp/libpsoap_new.so psoap_b200/csrc/libpsoap_b200.so
