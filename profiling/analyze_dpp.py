"""CLI wrapper: ``python profiling/analyze_dpp.py --static trace-pp --dpp trace-dpp --plots plots``
(reference ``profiling/process_*.py`` family)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatronapp_b200.dpp.profiling import main  # noqa: E402

if __name__ == "__main__":
    main()
