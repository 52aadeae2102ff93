import logging  # `from ditk import logging` (lzero/model/common.py:17)
