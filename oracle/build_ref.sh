#!/usr/bin/env bash
# Builds oracle/_ref/libfmref.so: the UNMODIFIED search code of the reference (/root/reference) behind the
# C interface of oracle/ref_shim.cpp.  TEST INFRASTRUCTURE ONLY.
#
# The reference is a header-only C++23 library that uses "deducing this" (needs GCC >= 14); this image has
# g++ 13.3.  The headers are therefore copied to a scratch directory OUTSIDE the repo, rewritten mechanically
# (no semantic change: `this auto&& self` in serialize()/operator[] becomes a plain `self` alias) and compiled
# from there.  Only the resulting .so lands in oracle/_ref/ (git-ignored).  Un-vendored third-party headers
# (libsais 2.10.4, mmser 1.0.1; cpmpack.json of the reference) are replaced by the stand-ins in oracle/shim/.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${FMREF_SRC:-/root/reference/src/fmindex-collection}"
if [ ! -d "$ref" ]; then echo "reference sources not found at $ref" >&2; exit 3; fi
tmp="$(mktemp -d /tmp/fmref.XXXXXX)"
trap 'rm -rf "$tmp"' EXIT
cp -r "$ref" "$tmp/fmindex-collection"
chmod -R u+w "$tmp"
cd "$tmp/fmindex-collection"
# 1. serialize/save/load members: (this auto&& self, Archive& ar) -> (Archive& ar) { auto& self = *this;
find . -name '*.h' -print0 | xargs -0 sed -E -i \
  's/\(this auto&& self, Archive& ar\)( const)? \{/(Archive\& ar)\1 { auto\& self = *this;/'
# 2. VectorBool::operator[](this auto& self, size_t idx)
sed -E -i 's/auto operator\[\]\(this auto& self, size_t idx\) \{/auto operator[](size_t idx) -> Proxy { return at(idx); }\n    auto operator[](size_t idx) const -> bool { auto\& self = *this;/' VectorBool.h
# 3. recursive lambdas `[&](this auto&& self, ...)` (not on the search path, but pulled in by includes):
#    explicit self parameter instead
for f in string/PairedFlattenedBitvectors2L.h string/PairedFlattenedBitvectors2LPartialSymb.h; do
  sed -E -i 's/\[&\]\(this auto&& self, /[\&](auto\&\& self, /; s/([^_a-zA-Z])self\(l_b1, b1,/\1self(self, l_b1, b1,/; s/([^_a-zA-Z])self\(b1, r_b1,/\1self(self, b1, r_b1,/; s/^([[:space:]]*)rec \($/\1rec (rec,/' "$f"
done
mkdir -p "$here/_ref"
g++ -std=c++23 -O3 -march=x86-64-v3 -DNDEBUG -fPIC -shared -pthread \
    -I "$here/shim" -I "$tmp" -I "$here" \
    "$here/ref_shim.cpp" -o "$here/_ref/libfmref.so"
echo "built $here/_ref/libfmref.so"
