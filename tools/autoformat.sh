#!/bin/bash
# Repository checks run before a commit (reference tools/autoformat.sh role; no formatter is vendored here, so this
# verifies instead of rewriting): syntax / whitespace / line-length lint and the undefined-name check.
set -e
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
cd "$ROOT"
python tools/linter.py "$@"
python tools/check_names.py
