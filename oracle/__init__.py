"""ORACLE — CPU restatements of the reference hot path. Test infrastructure only: nothing under
bdbnn_b200/, models/, kurtosis.py or utils/ may import this package (tests enforce it)."""
