"""Drop-in name: `import torbi` resolves to torbi_b200.

The reference's call sites (`torbi.from_probabilities`, `torbi.from_file`,
`torbi.from_file_to_file`, `torbi.from_files_to_files`, `torbi.data.loader`,
`python -m torbi ...`; /root/reference/torbi/__init__.py:72-78) run unchanged:
this module replaces itself and every `torbi.<submodule>` entry in sys.modules
by the corresponding torbi_b200 module, so `torbi.core is torbi_b200.core`,
configuration constants set through either name are the same objects, and
there is exactly one copy of the ctypes handle to libtorbi_b200.so.
"""
import sys

import torbi_b200

for _name, _module in list(sys.modules.items()):
    if _name == 'torbi_b200' or _name.startswith('torbi_b200.'):
        sys.modules['torbi' + _name[len('torbi_b200'):]] = _module
