"""TEST INFRASTRUCTURE ONLY - stand-in for pykeops==2.2.2 (absent, no network).

Only used by oracle/make_golden.py to execute the *real* reference
(/root/reference/src/losses/focus.py) in this container.  Never imported by the
product package.
"""
