# -*- coding: utf-8 -*-
"""
oracle.scenes_ext -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Scenes for the mesh optics and the plasma sources (filled in as those rows of
the scope table come up).
"""
EXTRA = {}
