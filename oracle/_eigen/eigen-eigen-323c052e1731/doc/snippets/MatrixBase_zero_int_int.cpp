cout << MatrixXi::Zero(2,3) << endl;
