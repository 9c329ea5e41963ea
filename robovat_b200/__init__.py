"""robovat_b200: B200-native PushEnv substep path behind robovat's API.

Importing the package does not load CUDA; `robovat_b200._capi.load()` does, and it
raises when libb2s.so is missing (there is no CPU fallback).
"""
__version__ = '0.1.0'
