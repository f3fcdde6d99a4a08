#!/usr/bin/env bash
# Regenerates the polygon known-answer vectors of tests/golden/kat.json from the REFERENCE's own Zig code and
# compares them with the committed (oracle-generated) ones.  This is the one command that turns "parity
# unpinned" into "pinned" for the triangulation; it cannot run in this repository's build image (no Zig
# toolchain, no network), so it is shipped untested -- run it on any machine with Zig >= 0.14.0-dev.2577
# (the reference's minimum_zig_version, build.zig.zon:18):
#
#     scripts/zig_golden.sh /path/to/myrenderer        # reference checkout (default /root/reference)
#
# What it does (nothing is copied into this repository; the work directory is a mktemp):
#   1. copies Polygon/Triangulation.zig and Polygon/unirand.zig from the reference checkout;
#   2. appends an explicit-pair constructor to unirand.zig and points Triangulation.zig:483 at it -- the only
#      source edit: `unirand_seed` draws from std.crypto.random (unirand.zig:31), which no test can reproduce;
#   3. builds scripts/zig_golden/main.zig against them (ReleaseSafe: a `.?` on null panics instead of being UB;
#      polygon1 and polygon2 never hit one) and runs it with stderr discarded (the reference prints from its hot
#      loops, Triangulation.zig:142,181,194,...; the prints do not affect results);
#   4. compares the emitted point ids with tests/golden/kat.json for every (offset, prime) pair.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
command -v zig >/dev/null || { echo "zig not found (need >= 0.14.0-dev.2577)"; exit 2; }
W="$(mktemp -d)"
trap 'rm -rf "$W"' EXIT
cp "$REF/Polygon/Triangulation.zig" "$REF/Polygon/unirand.zig" "$W/"
cp "$HERE/zig_golden/main.zig" "$W/"
cat >> "$W/unirand.zig" <<'ZIG'

// --- appended by scripts/zig_golden.sh: explicit (offset, prime) instead of std.crypto.random ---
pub var forced_offset: u32 = 0;
pub var forced_prime: u32 = 1;
pub fn unirand_explicit(top: u32) Unirand {
    return Unirand{ .at = 0, .top = top, .offset = forced_offset, .prime = forced_prime };
}
ZIG
sed -i 's/unirand\.unirand_seed(@intCast(points\.len))/unirand.unirand_explicit(@intCast(points.len))/' "$W/Triangulation.zig"
grep -q 'unirand_explicit' "$W/Triangulation.zig" || { echo "could not patch Triangulation.zig:483"; exit 3; }
(cd "$W" && zig build-exe -O ReleaseSafe main.zig)
"$W/main" 2>/dev/null > "$W/zig_kat.jsonl"
python3 "$HERE/zig_golden_compare.py" "$W/zig_kat.jsonl" "$HERE/../tests/golden/kat.json"
