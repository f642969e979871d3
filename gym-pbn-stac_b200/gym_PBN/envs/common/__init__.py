"""Truth-table PBN / PBCN cores (device-backed)."""
