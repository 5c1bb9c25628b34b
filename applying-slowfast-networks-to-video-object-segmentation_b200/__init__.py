"""B200-native hot path of SlowFast-VOS.  Import through the alias package ``sfvos_b200``."""
