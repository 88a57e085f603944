# Package directory of the B200-native VUnet hot path.  Import it as ``bdvs_b200`` (the hyphenated
# directory name is not a Python identifier; ``bdvs_b200/__init__.py`` at the repo root aliases it).
