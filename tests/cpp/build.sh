#!/usr/bin/env bash
# Builds the C++ tests of the host-side shim into tests/cpp/_bin/ (git-ignored, travels to the GPU box):
#   shim_test    include/fmb200/*.hpp against the CPU oracle (always)
#   dropin_test  the same search calls on the reference's own fmc::BiFMIndex (CPU) and on fmb200::attach(index) (GPU);
#                needs /root/reference (scratch copy + mechanical GCC-13 rewrite exactly as oracle/build_ref.sh)
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
lib="$root/fmindex-collection_b200"
mkdir -p "$here/_bin"
gcc -O2 -std=c11 -fPIC -c "$root/oracle/fm_oracle.c" -o "$here/_bin/fm_oracle.o"
g++ -std=c++20 -O2 -Wall -Wextra -I "$root/include" -I "$root/oracle" "$here/shim_test.cpp" "$here/_bin/fm_oracle.o" \
    -L "$lib" -lfmb200 -Wl,-rpath,'$ORIGIN/../../../fmindex-collection_b200' -pthread -o "$here/_bin/shim_test"
echo "built $here/_bin/shim_test"
ref="${FMREF_SRC:-/root/reference/src/fmindex-collection}"
if [ -d "$ref" ]; then
  tmp="$(mktemp -d /tmp/fmdropin.XXXXXX)"
  trap 'rm -rf "$tmp"' EXIT
  cp -r "$ref" "$tmp/fmindex-collection"
  chmod -R u+w "$tmp"
  ( cd "$tmp/fmindex-collection"
    find . -name '*.h' -print0 | xargs -0 sed -E -i 's/\(this auto&& self, Archive& ar\)( const)? \{/(Archive\& ar)\1 { auto\& self = *this;/'
    sed -E -i 's/auto operator\[\]\(this auto& self, size_t idx\) \{/auto operator[](size_t idx) -> Proxy { return at(idx); }\n    auto operator[](size_t idx) const -> bool { auto\& self = *this;/' VectorBool.h
    for f in string/PairedFlattenedBitvectors2L.h string/PairedFlattenedBitvectors2LPartialSymb.h; do
      sed -E -i 's/\[&\]\(this auto&& self, /[\&](auto\&\& self, /; s/([^_a-zA-Z])self\(l_b1, b1,/\1self(self, l_b1, b1,/; s/([^_a-zA-Z])self\(b1, r_b1,/\1self(self, b1, r_b1,/; s/^([[:space:]]*)rec \($/\1rec (rec,/' "$f"
    done )
  g++ -std=c++23 -O2 -DNDEBUG -I "$root/include" -I "$root/oracle/shim" -I "$tmp" "$here/dropin_test.cpp" \
      -L "$lib" -lfmb200 -Wl,-rpath,'$ORIGIN/../../../fmindex-collection_b200' -pthread -o "$here/_bin/dropin_test"
  echo "built $here/_bin/dropin_test"
else
  echo "reference sources not found at $ref: dropin_test not rebuilt" >&2
fi
