"""`python -m torbi ...` = `python -m torbi_b200 ...` (same flags as
/root/reference/torbi/__main__.py:12-50)."""
from torbi_b200.__main__ import main

if __name__ == '__main__':
    main()
