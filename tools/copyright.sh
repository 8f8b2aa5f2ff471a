#!/bin/bash
# List source files whose first lines carry neither a module docstring nor a comment header (reference
# tools/copyright.sh role: header hygiene).  Exit 1 if any.
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
bad=0
while IFS= read -r f; do
  first=$(grep -v '^#!' "$f" | grep -v '^\s*$' | head -1)
  case "$first" in
    '"""'*|"'''"*|'#'*|'//'*|'/*'*) ;;
    *) echo "no header: ${f#$ROOT/}"; bad=1 ;;
  esac
done < <(find "$ROOT/megatronapp_b200" "$ROOT/tools" "$ROOT/tasks" -type f \( -name '*.py' -o -name '*.cu' -o -name '*.cuh' -o -name '*.cpp' -o -name '*.h' \) ! -name '__init__.py' ! -path '*/__pycache__/*')
exit $bad
