in0_liver.bed in1_brain.bed in2_hek.bed in3_heart.bed
