"""TEST INFRASTRUCTURE ONLY -- see oracle/README.md. Never imported by particle_simulator_b200."""
