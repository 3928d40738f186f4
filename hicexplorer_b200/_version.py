# tracks the reference release whose hicCorrectMatrix surface is mirrored (hicexplorer/_version.py:4)
__version__ = '3.7.6+b200.1'
