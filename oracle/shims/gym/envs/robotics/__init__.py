"""Import-only stubs so that `/root/reference/icem/environments/robotics.py` can be imported and its `cost_fn`s (pure
NumPy) called unbound; the Fetch simulator (MuJoCo) is absent. TEST INFRASTRUCTURE ONLY."""
