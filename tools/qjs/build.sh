#!/bin/bash
# Builds oracle/_ref/qjs_run against the Qt 6 libraries inside Nsight Compute (no Qt headers needed) plus two
# empty stand-ins for libglib / libgthread (Qt Core links them for an event dispatcher this tool never creates).
set -e
here="$(cd "$(dirname "$0")" && pwd)"; root="$(cd "$here/../.." && pwd)"
QT="$(ls -d /opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64 2>/dev/null | head -1)"
[ -f "$QT/libQt6Qml.so.6" ] || { echo "no Qt6Qml found (Nsight Compute host directory missing)"; exit 3; }
out="$root/oracle/_ref"; mkdir -p "$out"
echo 'void __amb_stub(void){}' > "$out/stub.c"
gcc -shared -fPIC -o "$out/libglib-2.0.so.0" "$out/stub.c" -Wl,-soname,libglib-2.0.so.0
gcc -shared -fPIC -o "$out/libgthread-2.0.so.0" "$out/stub.c" -Wl,-soname,libgthread-2.0.so.0
rm "$out/stub.c"
g++ -O1 -std=c++17 -o "$out/qjs_run" "$here/qjs_run.cpp" \
    -Wl,--unresolved-symbols=ignore-in-shared-libs -Wl,--no-as-needed "$out/libglib-2.0.so.0" "$out/libgthread-2.0.so.0" "$QT/libQt6Qml.so.6" "$QT/libQt6Core.so.6" \
    -Wl,--disable-new-dtags -Wl,-rpath,"$QT" -Wl,-rpath,"$out"
echo "built $out/qjs_run"
