#!/bin/bash
# A/B helper: run a command with another build of the library swapped in (AB_LIB, default tools/ab/libpsoap_old.so),
# then restore the current one.
set -e
cd "$(dirname "$0")/.."
cp psoap_b200/csrc/libpsoap_b200.so /tmp/libpsoap_new.so
cp "${AB_LIB:-tools/ab/libpsoap_old.so}" psoap_b200/csrc/libpsoap_b200.so
"$@" || true
cp /tmp/libpsoap_new.so psoap_b200/csrc/libpsoap_b200.so
