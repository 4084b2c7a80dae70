"""madaiemulator_b200 -- B200-native Gaussian-process likelihood / gradient / prediction engine
behind MADAIEmulator's libEmu hot path."""
__version__ = "0.1.0"
