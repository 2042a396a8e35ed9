"""TEST INFRASTRUCTURE — the CPU oracle. Import only from tests/, smoke() and bench.py's CPU legs."""
