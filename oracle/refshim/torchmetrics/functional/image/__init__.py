from .. import structural_similarity_index_measure, peak_signal_noise_ratio
